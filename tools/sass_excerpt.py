"""SASS evidence for the TMA + mbarrier DMMA GEMM: disassembles the shipped library (cuobjdump -sass) and prints the
producer / consumer / main-loop excerpts of gemm_tma_kernel<80, 64, 2, NN, 4 stages> as markdown.
Usage: python tools/sass_excerpt.py > profiles/r02_sass_gemm_tma.md"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'gpitch_b200', 'libgpitch_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
key = 'gemm_tma_kernelILi80ELi64ELi2ELb0ELb0ELb0ELi4E'
ins, on = [], False
for l in txt.split('\n'):
    if 'Function :' in l:
        on = key in l
        continue
    if on:
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\*', l)
        if m:
            ins.append((m.group(1), m.group(2).strip()))


def block(i0, i1):
    return '\n'.join('/*%s*/  %s' % it for it in ins[max(0, i0):i1])


def first(pat, start=0):
    for i in range(start, len(ins)):
        if re.search(pat, ins[i][1]):
            return i
    return -1


print('# SASS of `gpx::gemm_tma_kernel<80, 64, 2, NN, 4 stages>` (dense / TRMM launch of the C3 step)\n')
print('`cuobjdump -sass gpitch_b200/libgpitch_b200.so` (sm_100a, built by `python -m gpitch_b200.build`), distilled by '
      '`tools/sass_excerpt.py`; instruction encodings stripped.\n')
c = collections.Counter((t.split()[1] if t.startswith('@') else t.split()[0]) for _, t in ins)
want = ['DMMA.8x8x4', 'LDS.64', 'UTMALDG.3D', 'UTMALDG.4D', 'SYNCS.EXCH.64', 'SYNCS.ARRIVE.TRANS64', 'SYNCS.ARRIVE.TRANS64.A1T0',
        'SYNCS.PHASECHK.TRANS64.TRYWAIT', 'BAR.SYNC.DEFER_BLOCKING', 'LDGSTS.E.BYPASS.128', 'LDGSTS']
print('Static counts: ' + ', '.join('`%s` x %d' % (k, c.get(k, 0)) for k in want) + ' (%d instructions in total).\n' % len(ins))
print('No `LDGSTS` (cp.async) and a single `BAR.SYNC` (publishing the mbarrier initialisation): operand tiles arrive by '
      '`UTMALDG` (cp.async.bulk.tensor), stage hand-over is `SYNCS` (mbarrier) only.\n')
i = first(r'SYNCS\.EXCH')
print('## mbarrier initialisation\n\n```\n' + block(i, i + 9) + '\n```\n')
i = first(r'SYNCS\.ARRIVE\.TRANS64 RZ')
j = first(r'UTMALDG\.4D', i)
print('## producer (one elected lane): expect_tx, A tile = one 3-d box (128-byte swizzle), B tile = ONE 4-d box '
      '{8, 16, BN/8, 1} (64-byte swizzle)\n\n```\n' + block(i, j + 2) + '\n```\n')
i = first(r'SYNCS\.PHASECHK\.TRANS64\.TRYWAIT')
print('## consumer: wait on the stage\'s `full` mbarrier\n\n```\n' + block(i, i + 4) + '\n```\n')
# densest DMMA window = the full-tile main loop
d = [k for k, (_, t) in enumerate(ins) if t.startswith('DMMA')]
best = max(range(len(d) - 20), key=lambda k: -(d[k + 20] - d[k]))
print('## main loop (excerpt of the fully live k-tile): `LDS.64` fragment reads from the swizzled stage feeding '
      '`DMMA.8x8x4`\n\n(the `NOP` after every `DMMA` is ptxas\'s own scheduling for this instruction on sm_100a; the register-resident peak '
      'kernel `gpx_dmma_peak` has it too)\n\n```\n' + block(d[best] - 14, d[best + 20] + 2) + '\n```\n')
i = first(r'SYNCS\.ARRIVE\.TRANS64\.A1T0')
j = first(r'SYNCS\.ARRIVE\.TRANS64 RZ', i)
print('## consumer release (`empty` arrive) and re-arm of the stage released one k-tile earlier (wait `empty`, expect_tx)\n\n```\n'
      + block(i - 1, j + 1) + '\n```')


# ---- split-K instantiation: cluster barrier + distributed-shared-memory loads of the partial accumulators
key2 = 'gemm_tma_kernelILi80ELi80ELi2ELb0ELb1ELb0ELi3E'
ins2, on = [], False
for l in txt.split('\n'):
    if 'Function :' in l:
        on = key2 in l
        continue
    if on:
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\*', l)
        if m:
            ins2.append((m.group(1), m.group(2).strip()))
c2 = collections.Counter((t.split()[1] if t.startswith('@') else t.split()[0]) for _, t in ins2)
k = next((i for i, (_, t) in enumerate(ins2) if t.startswith('UCGABAR_ARV')), -1)
print('\n## split-K over a thread-block cluster (`gemm_tma_kernel<80, 80, 2, NT, 3 stages>`, the single-window `A A^T` launch)\n')
print('Static counts: `UCGABAR_ARV` x %d, `UCGABAR_WAIT` x %d (barrier.cluster arrive / wait), `LD.E.64` x %d (the leader\'s '
      '`ld.shared::cluster` reads of the other CTAs\' accumulators; the `PRMT ..., 0x654` before them is `mapa`).\n' % (
          c2.get('UCGABAR_ARV', 0), c2.get('UCGABAR_WAIT', 0), c2.get('LD.E.64', 0)))
if k >= 0:
    print('```\n' + '\n'.join('/*%s*/  %s' % it for it in ins2[k - 1:k + 14]) + '\n```')
