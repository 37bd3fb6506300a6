"""Runs the short legs of bench.py for the named workloads only (development helper): python tools/leg_bench.py c1 c2"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gpitch_b200 import _lib

torch.cuda.set_device(0)
peak = _lib.dmma_peak(3)
for name in sys.argv[1:]:
    if name == 'predict':
        r = bench.run_predict_leg(torch.device('cuda', 0), peak, 6553.3, 'reference', 32.0)
        print(name, 'value %.1f windows/s  %.3f ms/step' % (r['value'], r['ms_per_step']), r['sanity'])
        for k, e in sorted(r['entry_points'].items(), key=lambda kv: -kv[1]['ms']):
            print('   %-14s %8.3f ms  %3d launches  share %.3f  %8.2f %s  frac %.3f' % (
                k, e['ms'], e['launches'], e['share'], e['achieved'], e['unit'], e['frac']))
        continue
    r = bench.run_leg(name, torch.device('cuda', 0), peak, 6553.3, 'reference', 32.0)
    print(name, 'value %.1f evals/s  %.3f ms/step  graph=%s  serialised eager %.3f ms' % (
        r['value'], r['ms_per_step'], r['cuda_graph_replay'], r['serialised_eager_step_ms']))
    for k, e in sorted(r['entry_points'].items(), key=lambda kv: -kv[1]['ms']):
        print('   %-14s %8.3f ms  %3d launches  share %.3f  %8.2f %s  frac %.3f' % (
            k, e['ms'], e['launches'], e['share_of_serialised_step'], e['achieved'], e['unit'], e['frac']))
