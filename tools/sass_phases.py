"""Static SASS histogram of a kernel, cut at its BAR.SYNC instructions (one fully unrolled phase of the batch loop per
segment).  Usage: python tools/sass_phases.py <mangled-kernel-name-substring> [lib.so]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'multi-modal-image-fusion_b200', 'libmmif_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
blocks = re.split(r'\n\s*Function : ', sass)
for blk in blocks[1:]:
    name = blk.split('\n', 1)[0]
    if sys.argv[1] not in name:
        continue
    ins = []
    for l in blk.split('\n'):
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            txt = re.sub(r'^@!?U?P\d+\s+', '', m.group(2).strip())
            ins.append((int(m.group(1), 16), txt.split()[0].split('.')[0]))
    print(name, len(ins), 'instructions', '%.1f KB' % (len(ins) * 16 / 1024))
    seg, cur, start = [], collections.Counter(), 0
    for i, (pc, op) in enumerate(ins):
        cur[op] += 1
        if op == 'BAR':
            seg.append((start, i, cur)); cur = collections.Counter(); start = i + 1
    seg.append((start, len(ins), cur))
    for a, b, c in seg:
        n = sum(c.values())
        if n < 40:
            continue
        fma = 2 * (c['FFMA2'] + c['FMUL2'] + c['FADD2']) + c['FFMA'] + c['FMUL'] + c['FADD']
        print(f'[{a:5d},{b:5d}) n={n:5d} fma-pipe slots={fma:5d}  ' + ', '.join(f'{k}:{v}' for k, v in c.most_common(16)))
