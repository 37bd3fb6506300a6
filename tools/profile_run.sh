#!/bin/bash
# ncu launch list of the bench step (run under gpurun): 17 windows = one window chunk of the C3 workload;
# 3 warm-up + 1 timed step.
set -x
CMD="python bench.py --windows 16 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
