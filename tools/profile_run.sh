#!/bin/bash
# ncu evidence for the bench step (run under gpurun): launch list + one full capture of the dominant kernels.
set -x
CMD="python bench.py --windows 17 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 648 -c 260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 180 -c 8 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"build_kernel|grad_kernel" -s 12 -c 4 -o gpurun_out/prof_builder $CMD > gpurun_out/ncu_builder.log 2>&1
ls -la gpurun_out/
