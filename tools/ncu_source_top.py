"""Top stalled SASS instructions of an `ncu --page source --csv` dump (gz): samples, dominant stall reason, executed count.
usage: python tools/ncu_source_top.py profiles/r02_ncu_tri_source.csv.gz [N]"""
import csv, gzip, io, sys
rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][1][:120])
h = rows[1]
iS, iN, iX = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
st = [i for i, k in enumerate(h) if k.startswith('stall_') and 'Not Issued' not in k]
body = rows[2:]
tot = sum(int(r[iS]) for r in body)
print('total samples', tot, 'instructions', len(body))
agg = {}
for r in body:
    for i in st:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print('by reason:', ', '.join(f'{k[6:]} {100*v/tot:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda j: -int(body[j][iS]))[:N]
for j in sorted(order):
    r = body[j]
    top = max(st, key=lambda i: int(r[i]))
    print(f'{j:5d} {100*int(r[iS])/tot:5.1f}%  x{r[iX]:>9}  {h[top][6:]:<12} {r[iN].strip()[:100]}')
