"""Contiguous SASS blocks of a capture with equal execution counts: static size, dynamic instructions, FFMA2 / LDS / ALU mix."""
import csv, subprocess, sys, re
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr, data = rows[k], rows[k + 1:]
ix = {h: i for i, h in enumerate(hdr)}
iS, iI, iN = ix['Source'], ix['Instructions Executed'], ix['# Samples']
tot = sum(int(r[iI]) for r in data)
unit = float(sys.argv[2]) if len(sys.argv) > 2 else None
blocks = []
cur = None
for i, r in enumerate(data):
    c = int(r[iI])
    if cur and (c == cur['c'] or (min(c, cur['c']) > 0 and 0.97 < c / cur['c'] < 1.03)):
        cur['n'] += 1; cur['dyn'] += c; cur['rows'].append(r)
    else:
        cur = {'start': i, 'c': c, 'n': 1, 'dyn': c, 'rows': [r]}
        blocks.append(cur)
print(f'total dynamic {tot}')
for b in blocks:
    if b['dyn'] * 200 < tot:
        continue
    ops = {}
    for r in b['rows']:
        s = re.sub(r'^@!?U?P\d+\s+', '', r[iS].strip())
        op = s.split()[0].split('.')[0] if s else ''
        ops[op] = ops.get(op, 0) + 1
    samp = sum(int(r[iN]) for r in b['rows'])
    top = ' '.join(f'{o}:{v}' for o, v in sorted(ops.items(), key=lambda kv: -kv[1])[:9])
    per = f" per-unit {b['dyn'] / unit:7.1f}" if unit else ''
    print(f"[{b['start']:5d}+{b['n']:4d}] count {b['c']:9d} dyn {100 * b['dyn'] / tot:5.1f}%{per} samples {samp:6d} | {top}")
