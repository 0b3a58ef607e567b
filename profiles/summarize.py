#!/usr/bin/env python
"""Turn ncu outputs into the markdown summaries kept in profiles/.

    python profiles/summarize.py launches <launches.csv> "<title>" "<command>" > profiles/rNN_x_launches_summary.md
    python profiles/summarize.py full <report.ncu-rep> "<title>" > profiles/rNN_x_ncu_full_summary.md

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --csv`; `full` calls
`ncu -i <rep> --page raw --csv` and keeps the metrics the roofline discussion needs.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = OrderedDict([
    ('gpu__time_duration.sum', 'time us'),
    ('dram__bytes_read.sum', 'dram rd MB'),
    ('dram__bytes_write.sum', 'dram wr MB'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm %'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64 pipe %'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed', 'smem %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
])


def short(name):
    name = re.sub(r'\(.*', '', name)
    name = name.replace('void ', '').replace('lmc::', '')
    return name


def rows_of(text):
    lines = [ln for ln in text.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO('\n'.join(lines))))


def launches(path, title, command):
    rows = rows_of(open(path).read())
    hdr = rows[0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = OrderedDict()
    for r in rows[1:]:
        k = short(r[ik])
        ns = float(r[iv].replace(',', ''))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print('# %s\n' % title)
    print('Command: `%s`' % command)
    print('(cold-cache, serialised launches: compare shares, not absolutes)\n')
    print('| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|')
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.3f | %.1f | %.1f%% |' % (k, c, ns / 1e6, ns / c / 1e3, 100 * ns / tot))


def full(path, title):
    if path.endswith('.csv'):
        out = open(path).read()
    else:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = rows_of(out)
    hdr, units = rows[0], rows[1]
    ik = hdr.index('Kernel Name')
    cols = [(hdr.index(m), lab, m) for m, lab in KEEP.items() if m in hdr]
    seen = OrderedDict()
    for r in rows[2:]:
        k = short(r[ik]) + ' grid=' + r[hdr.index('launch__grid_size')] if 'launch__grid_size' in hdr else short(r[ik])
        seen.setdefault(k, r)      # first launch of every distinct (kernel, grid)
    print('# %s\n' % title)
    print('One launch per distinct (kernel, grid); `ncu --set full --clock-control none --import-source on`.\n')
    print('| kernel | ' + ' | '.join(lab for _, lab, _ in cols) + ' |')
    print('|---|' + '---|' * len(cols))
    for k, r in seen.items():
        vals = []
        for i, lab, m in cols:
            v = r[i].replace(',', '')
            try:
                f = float(v)
                u = units[i]
                if m.startswith('dram__bytes'):
                    f = f / 1e6 if u == 'byte' else (f * 1e3 if u == 'Gbyte' else (f / 1e3 if u == 'Kbyte' else f))
                if m == 'gpu__time_duration.sum':
                    f = f / 1e3 if u in ('ns', 'nsecond') else (f * 1e3 if u in ('ms', 'msecond') else f)
                vals.append('%.1f' % f if f != int(f) else '%d' % f)
            except ValueError:
                vals.append(v)
        print('| `%s` | ' % k + ' | '.join(vals) + ' |')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
