#!/usr/bin/env python
"""Per kernel: stall-reason totals and the hottest SASS instructions from `ncu --page source --csv`."""
import csv, gzip, io, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
txt = (gzip.open(path, 'rt') if path.endswith('.gz') else open(path)).read()
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == 'Kernel Name':
        cur = {'name': row[1], 'hdr': None, 'rows': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = row
    elif cur is not None:
        cur['rows'].append(row)
def num(x):
    try: return float(x.replace(',', ''))
    except ValueError: return 0.0
for b in blocks:
    h = b['hdr']
    isamp = h.index('# Samples')
    stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    tot = sum(num(r[isamp]) for r in b['rows']) or 1
    print('==', b['name'][:110], ' samples', int(tot), ' sass lines', len(b['rows']))
    st = sorted(((sum(num(r[i]) for r in b['rows']), h[i]) for i in stall), reverse=True)
    print('   stalls:', ', '.join('%s %.0f%%' % (n, 100 * v / tot) for v, n in st[:7]))
    def col(name):
        return h.index(name) if name in h else None
    ex_s, id_s = col('L1 Wavefronts Shared'), col('L1 Wavefronts Shared Ideal')
    ex_g, id_g = col('L2 Theoretical Sectors Global'), col('L2 Theoretical Sectors Global Ideal')
    if ex_s is not None:
        print('   smem wavefronts %.3g (ideal %.3g); global sectors %.3g (ideal %.3g)' % (
            sum(num(r[ex_s]) for r in b['rows']), sum(num(r[id_s]) for r in b['rows']),
            sum(num(r[ex_g]) for r in b['rows']), sum(num(r[id_g]) for r in b['rows'])))
    iex = h.index('Instructions Executed')
    print('   warp instructions executed: %.4g' % sum(num(r[iex]) for r in b['rows']))
    idx = sorted(range(len(b['rows'])), key=lambda i: -num(b['rows'][i][isamp]))[:top]
    for i in sorted(idx):
        r = b['rows'][i]
        rs = sorted(((num(r[j]), h[j][6:]) for j in stall), reverse=True)[:2]
        print('   %5d %5.1f%%  %-58s %s' % (i, 100 * num(r[isamp]) / tot, r[1].strip()[:58],
                                            ' '.join('%s=%d' % (n, v) for v, n in rs if v)))
