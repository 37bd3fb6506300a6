"""Distil `ncu --page raw --csv` exports into a markdown table (one row per profiled launch)."""
import csv, sys
cols = [('Kernel Name', 'kernel'), ('Grid Size', 'grid'), ('gpu__time_duration.sum', 'time'),
        ('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'DMMA pipe %'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 ALU pipe %'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'), ('launch__registers_per_thread', 'regs'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %')]
print('| ' + ' | '.join(c[1] for c in cols) + ' |')
print('|' + '---|' * len(cols))
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        out = []
        for c, _ in cols:
            if c not in idx:
                out.append('-'); continue
            v, u = r[idx[c]], units[idx[c]]
            if c == 'Kernel Name':
                v = v.replace('void ', '').replace('(GemmArgs)', '').replace('(KernArgs)', '').replace('(KernArgs, int)', '')[:44]
                out.append('`%s`' % v)
            elif c == 'Grid Size':
                out.append(v)
            else:
                try:
                    out.append('%.3g %s' % (float(v.replace(',', '')), u if u not in ('%', 'register/thread') else ''))
                except ValueError:
                    out.append(v)
        print('| ' + ' | '.join(out) + ' |')
