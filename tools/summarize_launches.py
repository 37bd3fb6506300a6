"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --windows 17 --steps 1 --warmup 3`
into a per-kernel share table of ONE step (the launches between the 3rd and 4th varexp launch = backward of a
step + forward of the next = the multiset of one step).  Usage: python tools/summarize_launches.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ki, vi, ui, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
recs = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', '')) * {'us': 1e-3, 'ns': 1e-6, 'ms': 1.0}.get(r[ui], 1.0)
    recs.append((r[ki], r[gi], v))
ve = [i for i, (n, g, v) in enumerate(recs) if 'varexp' in n]
if len(ve) >= 4:            # (the workload legs that follow the timed step launch more of them: take the 3rd -> 4th)
    seg = recs[ve[2]:ve[3]]
elif len(ve) >= 2:
    seg = recs[ve[-2]:ve[-1]]
else:                       # SGPR workloads have no quadrature launch: 3 warm-up + 1 timed step -> the last quarter
    seg = recs[-(len(recs) // 4):]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, g, v in seg:
    key = n.split('(')[0].replace('void ', '')[:64]
    if 'gemm_tma_kernel' in n:      # 1-d grid: n-tiles x m-tiles x batch
        gx = int(g.strip('()').split(',')[0])
        key = key.replace('gpx::_GLOBAL__N__', 'gpx::').split('gemm_tma_cu_')[-1] if '_GLOBAL__N__' in key else key
        key += '  [M x M x N, N=4000]' if gx >= 20000 else '  [M x M x M and smaller]'
    elif 'gemm_kernel' in n:
        gx = int(g.strip('()').split(',')[0])
        key += '  [M x M x N, N=4000]' if gx >= 16 else '  [M x M x M and smaller]'
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if k.startswith('gpx::'))
print('| kernel | launches | ms (serialised, cold cache) | share |')
print('|---|---:|---:|---:|')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.0005:
        continue
    print('| `%s` | %d | %.3f | %.1f %% |' % (k, v[0], v[1], 100 * v[1] / tot))
print('\nlaunches in one step of one window chunk: %d; sum of kernel times %.2f ms; gpx:: kernels %.1f %% of it' % (
    len(seg), tot, 100 * ours / tot))
