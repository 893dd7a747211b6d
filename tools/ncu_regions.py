"""Samples of a kernel aggregated over SASS index ranges: python tools/ncu_regions.py rep kernel-substr b0,b1,b2,..."""
import csv, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
bounds = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else []
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
blocks = []; cur = None
for ln in txt.splitlines():
    if ln.startswith('"Kernel Name"'):
        cur = {'name': ln, 'rows': []}; blocks.append(cur)
    elif ln.startswith('"Address"'):
        cur['hdr'] = next(csv.reader([ln]))
    elif cur is not None and 'hdr' in cur:
        cur['rows'].append(next(csv.reader([ln])))
blk = [b for b in blocks if sub in b['name']][0]
h = {k: i for i, k in enumerate(blk['hdr'])}
rows = blk['rows']
stall_cols = [k for k in blk['hdr'] if k.startswith('stall_') and 'Not Issued' not in k]
bounds = [0] + bounds + [len(rows)]
tot = sum(int(r[h['# Samples']] or 0) for r in rows)
for a, b in zip(bounds[:-1], bounds[1:]):
    rr = rows[a:b]
    s = sum(int(r[h['# Samples']] or 0) for r in rr)
    ex = sum(int(r[h['Instructions Executed']] or 0) for r in rr)
    agg = {k[6:]: sum(int(r[h[k]] or 0) for r in rr) for k in stall_cols}
    agg = {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1])[:5] if v}
    print(f'[{a:5d},{b:5d}) samples {s:6d} ({100*s/max(tot,1):5.1f}%) warp-instr {ex:10d}  {agg}')
