"""Print the per-launch durations of the n-th forward found in an ncu launch list (gpu__time_duration.sum csv)."""
import csv, re, sys
path, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
L = []
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
    L.append((r[ki], v))
starts = [i for i, (n, v) in enumerate(L) if 'ncdhw_to_blocked' in n][::2]
s = starts[which]; e = starts[which + 1] if which + 1 < len(starts) else len(L)
tot = 0
for i, (n, v) in enumerate(L[s:e]):
    m = re.search(r'conv3d_tc_kernel<(.*?)>', n)
    tot += v
    print(i, m.group(1) if m else n[:50], round(v, 3))
print('total', round(tot, 3))
