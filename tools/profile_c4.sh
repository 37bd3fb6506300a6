cd gpurun_out
python ../tools/kernel_bench_c4.py 1 > c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"build_kernel_sum|grad_lag_bin" -c 4 -o c4prof python ../tools/kernel_bench_c4.py 1 > ncu_c4.log 2>&1
ncu -i c4prof.ncu-rep --page raw --csv > c4_raw.csv 2>/dev/null
ncu -i c4prof.ncu-rep --page source --csv 2>/dev/null | gzip > c4_source.csv.gz; rm -f c4prof.ncu-rep; ls -la c4*
