"""Print the per-layer table of one or more bench JSON lines: python tools/show_bench.py a.json b.json ..."""
import json
import sys

NAMES = {0: 'dres0.0', 1: 'dres0.2', 2: 'dres1.0', 3: 'dres1.2', 4: 'conv1', 5: 'conv2', 6: 'conv3', 7: 'conv4', 8: 'conv5', 9: 'conv6',
         22: 'classif.0', 25: 'head', -1: 'cv', -2: 'softargmin'}


def load(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        return None


rows = {p.split('/')[-1].replace('.json', '')[-12:]: load(p) for p in sys.argv[1:]}
for k, j in rows.items():
    if j:
        print(f"{k:14s} ms/step {j['ms_per_step']:.2f}  value {j['value']:.1f}  e2e {j['e2e']['value']:.1f}  clk {j['clocks']['sm_mhz']}  conv_ms {j['roofline']['conv_ms_per_step']:.2f}")
print('layer          ', ' '.join(f'{k[-8:]:>8}' for k in rows))
for k in [-2, -1] + list(range(0, 10)) + [22, 25]:
    print(f'{k:3d} {NAMES.get(k, ""):10s}', ' '.join(f"{(rows[c]['ms_by_layer'].get(str(k), 0) if rows[c] else 0):8.3f}" for c in rows))
tot = {c: (sum(v for kk, v in rows[c]['ms_by_layer'].items()) if rows[c] else 0) for c in rows}
print('sum            ', ' '.join(f'{tot[c]:8.2f}' for c in rows))
