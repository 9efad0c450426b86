"""Stall table of a warp-specialised kernel capture, cut at the mbarrier waits / arrives / named barriers."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr, data = rows[k], rows[k + 1:]
ix = {h: i for i, h in enumerate(hdr)}
iS, iN, iI = ix['Source'], ix['# Samples'], ix['Instructions Executed']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iN]) for r in data) or 1
totI = sum(int(r[iI]) for r in data) or 1
print(f'instructions {len(data)}  samples {tot}  warp instructions {totI}')
start = 0
for i in range(len(data) + 1):
    cut = i == len(data) or any(t in data[i][iS] for t in ('SYNCS.ARRIVE', 'BAR.SYNC', 'TRYWAIT', 'EXIT'))
    if not cut:
        continue
    stop = min(i + 1, len(data))
    seg = data[start:stop]
    samp = sum(int(r[iN]) for r in seg)
    inst = sum(int(r[iI]) for r in seg)
    if samp * 100 > tot:
        st = {s: sum(int(r[ix[s]]) for r in seg) for s in stalls}
        ts = sorted(st.items(), key=lambda kv: -kv[1])[:5]
        print(f'[{start:5d},{stop:5d}) n {stop - start:4d} samples {100 * samp / tot:5.1f}% inst {100 * inst / totI:5.1f}% | ' +
              ' '.join(f'{s[6:]}:{100 * v / tot:.1f}' for s, v in ts) + ' | ends: ' + data[stop - 1][iS].strip()[:50])
    start = stop
