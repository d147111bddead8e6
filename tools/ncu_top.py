"""Summarise the `ncu --page source --csv` export: top stall locations (SASS) with reasons."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
iS = h.index('# Samples'); isrc = h.index('Source')
data = [r for r in rows[2:] if len(r) == len(h) and r[iS].isdigit()]
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[iS]) for r in data)
print('rows', len(data), 'total samples', tot)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:n]
for i in sorted(top):
    r = data[i]
    st = {h[c]: int(r[c]) for c in stall_cols if int(r[c]) > 0}
    st = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(i, r[iS], '%.1f%%' % (100.0 * int(r[iS]) / tot), r[isrc].strip()[:70], st)
