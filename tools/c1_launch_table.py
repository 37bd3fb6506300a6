"""profiles/r02_c1_single_window.md from the ncu launch list of tools/c1_launch_list.py:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file c1.csv python tools/c1_launch_list.py
   python tools/c1_launch_table.py c1.csv > profiles/r02_c1_single_window.md"""
import collections, csv, sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
r = list(csv.DictReader(lines))
starts = [i for i, x in enumerate(r) if 'sgpr_prep_kernel' in x['Kernel Name']]
seg = r[starts[-1]:]
agg, tot = collections.OrderedDict(), 0.0
for x in seg:
    v = float(x['Metric Value'].replace(',', '')) / 1000
    tot += v
    k = x['Kernel Name'].split('(')[0].replace('void ', '').replace('gpx::<unnamed>::', 'gpx::')[:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
print('# One SGPRSS evaluation of a single configs[0] window (N = 1600, M = 200, P = 3, Q = 10), by kernel -- round 2, final state\n')
print('`tools/c1_launch_list.py` (three `BatchedSGPR.bound` calls = three `gpx_sgpr_bound` C calls) under')
print('`ncu --metrics gpu__time_duration.sum --clock-control none`; the table is the last call (`tools/c1_launch_table.py`).  Times under')
print('ncu are cold-cache and serialised; the same sequence replayed as one CUDA graph takes 0.61 ms (`bench.py` `workloads.c1`).\n')
print('| kernel | launches | us (serialised) | share |\n|---|---:|---:|---:|')
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('| `%s` | %d | %.1f | %.1f %% |' % (k, c, v, 100 * v / tot))
print("\n%d launches, %.0f us in total (1400 us before this round's single-window work: 33 us diagonal blocks, a 123 us `A A^T`" % (len(seg), tot))
print('SYRK on 6 CTAs, 25 us single-CTA column statistics, a 62 us lag-histogram pass).  The two Cholesky + inverse chains')
print('(`diag_block_kernel`, `zero_upper_kernel` and the 64-wide `gemm_tma_kernel<80, 64, 4, ...>` panel / update / inverse products) are')
print('the largest group; the `A A^T` SYRK (`<80, 80, 2, 0, 1, 0, 3>`) and the M x M x M products run split-K over thread-block clusters.')
