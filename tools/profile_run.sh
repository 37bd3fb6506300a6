#!/bin/bash
# ncu evidence for the bench step (run under gpurun): full launch list of a 1-chunk run (17 windows = one window
# chunk of the C3 workload; 3 warm-up + 1 timed step), then full captures of the dominant kernels.
set -x
CMD="python bench.py --windows 17 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ls -la gpurun_out/
