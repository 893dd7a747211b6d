"""Top stall sites of a kernel from an .ncu-rep (SASS page): python tools/ncu_hot.py rep [topN] [kernel-substr]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25; sub = sys.argv[3] if len(sys.argv) > 3 else ''
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
lines = txt.splitlines()
# find header line
blocks = []; cur = None
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        cur = {'name': ln, 'rows': []}; blocks.append(cur)
    elif ln.startswith('"Address"'):
        cur['hdr'] = next(csv.reader([ln]))
    elif cur is not None and 'hdr' in cur:
        cur['rows'].append(next(csv.reader([ln])))
for blk in [b for b in blocks if sub in b['name']][:1]:
    h = {k: i for i, k in enumerate(blk['hdr'])}
    rows = blk['rows']
    tot = sum(int(r[h['# Samples']] or 0) for r in rows)
    print(blk['name'][:120], 'total samples', tot, 'instrs', len(rows))
    stall_cols = [k for k in blk['hdr'] if k.startswith('stall_') and 'Not Issued' not in k]
    agg = {k: sum(int(r[h[k]] or 0) for r in rows) for k in stall_cols}
    print('stall totals:', {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][h['# Samples']] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        st = {k[6:]: int(r[h[k]] or 0) for k in stall_cols if int(r[h[k]] or 0)}
        print(f"{i:5d} {int(r[h['# Samples']]):6d} {r[h['Instructions Executed']]:>9s} {r[h['Source']][:70]:70s} {st}")
