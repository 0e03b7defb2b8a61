#!/bin/bash
# timing ablations of the tensor-core conv kernel (results are numerically meaningless; only per-layer ms matter)
for d in "$@"; do
  IDISP_TC_DBG=$d timeout 120 python bench.py --steps 2 --warmup 3 --precision bf16 --no-cpu-baseline 2>/dev/null > /tmp/dbg_$d.json
  python - "$d" <<'PY'
import json,sys
d=sys.argv[1]
try:
    j=json.load(open(f'/tmp/dbg_{d}.json')); m=j["ms_by_layer"]
    print("dbg", d, "step", round(j["ms_per_step"],2), {k: round(m[k],3) for k in ["0","1","4","5","6","8","9","25"]})
except Exception as e:
    print("dbg", d, "failed", e)
PY
done
