#!/bin/bash
# ncu evidence for the bench step (run under gpurun): full launch list of a 1-chunk run (17 windows = one window
# chunk of the C3 workload; 3 warm-up + 1 timed step), then full captures of the dominant kernels.
set -x
CMD="python bench.py --windows 17 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel" -s 250 -c 14 -o gpurun_out/prof_step_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"build_kernel|grad_kernel|varexp|colstats" -s 20 -c 8 -o gpurun_out/prof_step_builder $CMD > gpurun_out/ncu_builder.log 2>&1
ls -la gpurun_out/
