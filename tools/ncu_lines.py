"""Summarise `ncu --page source --csv --print-source cuda,sass`: stall samples per CUDA source line of one file
(SASS rows are attributed to the preceding source-line row)."""
import csv
import sys

path, fname = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.008
rows = list(csv.reader(open(path)))
start = [i for i, r in enumerate(rows) if len(r) >= 2 and r[0] == 'File Path' and r[1].endswith(fname)][0]
h = rows[start + 2]
iS, iI = h.index('# Samples'), h.index('Instructions Executed')
iW = h.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in h else 10 ** 6
iWi = h.index('L1 Wavefronts Shared Ideal') if 'L1 Wavefronts Shared Ideal' in h else 10 ** 6
stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
cur, src, agg = None, {}, {}
for r in rows[start + 3:]:
    if r and r[0] == 'File Path':
        break
    if r and r[0].isdigit():
        cur = int(r[0])
        src[cur] = ','.join(r[1:-60])[:80] if len(r) > 62 else r[1][:80]
        continue
    if cur is None or len(r) < 40 or r[0] != '' or not r[2].startswith('0x'):
        continue
    a = agg.setdefault(cur, {'s': 0, 'i': 0, 'w': 0, 'wi': 0, 'st': {}})
    g = lambda i: int(r[i]) if i < len(r) and r[i].isdigit() else 0
    a['s'] += g(iS); a['i'] += g(iI); a['w'] += g(iW); a['wi'] += g(iWi)
    for c in stall:
        if g(c):
            a['st'][h[c]] = a['st'].get(h[c], 0) + g(c)
tot = sum(a['s'] for a in agg.values())
print('total samples', tot)
for ln in sorted(agg):
    a = agg[ln]
    if a['s'] >= thr * tot:
        st = sorted(a['st'].items(), key=lambda x: -x[1])[:3]
        print('%5d %6d %5.1f%% inst %10d smem-wf %9d (ideal %9d) %s | %s' % (ln, a['s'], 100.0 * a['s'] / tot, a['i'], a['w'], a['wi'], st, src.get(ln, '')))
