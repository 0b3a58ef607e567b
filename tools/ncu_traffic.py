#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --page raw --csv` dumps: DRAM bytes (read + write) per RHS pair
and kernel family.   python tools/ncu_traffic.py E=<raw.csv>:<pairs> D=<raw.csv>:<pairs> ..."""
import csv, io, json, os, sys
FAM = [('to_grid', 'to_grid'), ('from_grid', 'from_grid'), ('fused_lines', 'mix'), ('fused_col512', 'mix'),
       ('rows512_fwd', 'fft_fwd_contig'), ('rows512_inv', 'fft_inv_contig'),
       ('fft_rows_T_kernel<0>', 'fft_fwd_contig'), ('fft_rows_T_fwd_pipe', 'fft_fwd_contig'), ('fft_rows_T_kernel<1>', 'fft_inv_contig'),
       ('fft_pass_kernel<1, 0>', 'fft_fwd_strided'), ('fft_pass_kernel<1, 1>', 'fft_inv_strided'),
       ('fft_pass_kernel<0, 0>', 'fft_fwd_contig'), ('fft_pass_kernel<0, 1>', 'fft_inv_contig'),
       ('permute_cols', 'other')]
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, 'profiles', 'ncu_traffic.json')
out = json.load(open(path)) if os.path.exists(path) else {}
src = out.setdefault('_sources', {})
for arg in sys.argv[1:]:
    wl, rest = arg.split('=')
    f, pairs = rest.rsplit(':', 1)
    lines = [ln for ln in open(f).read().splitlines() if ln.startswith('"')]
    rows = list(csv.reader(io.StringIO('\n'.join(lines))))
    hdr, units = rows[0], rows[1]
    ik, ir, iw = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    acc = {}
    for r in rows[2:]:
        for key, fam in FAM:
            if key in r[ik]:
                b = float(r[ir].replace(',', '')) * SCALE[units[ir]] + float(r[iw].replace(',', '')) * SCALE[units[iw]]
                acc[fam] = acc.get(fam, 0.0) + b
                break
    out[wl] = {k: v / int(pairs) for k, v in acc.items()}
    src[wl] = '%s (%s pairs in the captured block product)' % (os.path.basename(f), pairs)
    print(wl, {k: round(v / int(pairs) / 1e6, 2) for k, v in acc.items()}, 'MB per pair')
out['_note'] = 'dram__bytes_read.sum + dram__bytes_write.sum per RHS pair and kernel family, summed over the launches of one block product under ncu --set full'
json.dump(out, open(path, 'w'), indent=1)
