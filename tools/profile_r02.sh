#!/bin/bash
# Round-2 ncu evidence (run under gpurun).  --set full captures of the GEMM launch types (TRMM, SYRK, dense; each twice:
# warm-up + timed launch) and of the builder / gradient kernels in isolation, exported to CSV on the box (the .ncu-rep
# files exceed what gpurun copies back), plus one capture with source correlation of the TRMM launch for the SASS excerpt.
# Every profiled command first runs plain.  The launch list of the bench step comes from tools/profile_run.sh.
set -x
cd gpurun_out
python ../tools/gemm_bench.py 1 0,3,4 > gb_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:gemm_tma -c 6 -o r02_gemm python ../tools/gemm_bench.py 1 0,3,4 > ncu_gemm.log 2>&1
ncu -i r02_gemm.ncu-rep --page raw --csv > r02_gemm_raw.csv 2>/dev/null; rm -f r02_gemm.ncu-rep
python ../tools/kernel_bench.py 1 0,2,5,7,9 > kb_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"build_kernel|grad_kernel|grad_lag" -c 16 -o r02_kern python ../tools/kernel_bench.py 1 0,2,5,7,9 > ncu_kern.log 2>&1
ncu -i r02_kern.ncu-rep --page raw --csv > r02_kern_raw.csv 2>/dev/null; rm -f r02_kern.ncu-rep
python ../tools/gemm_bench.py 1 0 > gb_plain0.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 1 -c 1 -o r02_trmm_src python ../tools/gemm_bench.py 1 0 > ncu_trmm.log 2>&1
ncu -i r02_trmm_src.ncu-rep --page source --csv 2>/dev/null | gzip > r02_trmm_source.csv.gz; rm -f r02_trmm_src.ncu-rep
ls -la
