"""Per-phase breakdown of an `ncu --set full --import-source on` capture of a barrier-phased kernel.

    python tools/ncu_phases.py capture.ncu-rep [--split IDX ...]

The SASS source page is cut at every BAR.SYNC (and at the extra instruction indices given with
--split); for each piece: share of the warp-stall samples (~time), share of executed instructions,
top opcodes, top stall reasons.  `--dump A B` prints the SASS lines [A, B) with their counts.
"""
import csv
import re
import subprocess
import sys


def load(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    k = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    return rows[k], rows[k + 1:]


def main():
    rep = sys.argv[1]
    args = sys.argv[2:]
    hdr, data = load(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    iS, iN, iI = ix['Source'], ix['# Samples'], ix['Instructions Executed']
    if args and args[0] == '--dump':
        a, b = int(args[1]), int(args[2])
        for k in range(a, b):
            print(k, data[k][iI], data[k][iN], data[k][iS].strip()[:110])
        return
    splits = set(int(v) for v in args[1:]) if args and args[0] == '--split' else set()
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[iN]) for r in data) or 1
    totI = sum(int(r[iI]) for r in data) or 1
    print(f'total samples {tot}  warp instructions {totI}')
    start = 0
    for k in range(len(data) + 1):
        end = k == len(data) or 'BAR.SYNC' in data[k][iS] or k in splits
        if not end:
            continue
        stop = min(k + 1, len(data))
        seg = data[start:stop]
        samp = sum(int(r[iN]) for r in seg)
        inst = sum(int(r[iI]) for r in seg)
        if samp * 200 > tot or inst * 200 > totI:
            ops = {}
            for r in seg:
                s = re.sub(r'^@!?U?P\d+\s+', '', r[iS].strip())
                op = s.split()[0].split('.')[0] if s else ''
                ops[op] = ops.get(op, 0) + int(r[iI])
            st = {s: sum(int(r[ix[s]]) for r in seg) for s in stalls}
            top = sorted(ops.items(), key=lambda kv: -kv[1])[:10]
            ts = sorted(st.items(), key=lambda kv: -kv[1])[:5]
            print(f'[{start:5d},{stop:5d}) static {stop - start:5d}  samples {100 * samp / tot:5.1f}%  inst {100 * inst / totI:5.1f}%')
            print('      ops  ', ' '.join(f'{o}:{100 * v / totI:.2f}' for o, v in top))
            print('      stall', ' '.join(f'{s[6:]}:{100 * v / tot:.1f}' for s, v in ts))
        start = stop


if __name__ == '__main__':
    main()
