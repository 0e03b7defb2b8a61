set -x
mkdir -p gpurun_out
export IDISP_BENCH_SKIP_REFGPU=1
export IDISP_BENCH_SKIP_LIVE=1
for d in 0 4096 0 4096; do
  IDISP_TC_DBG=$d timeout 200 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r02y_$d.json 2>/dev/null
  python - $d <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r02y_{sys.argv[1]}.json') if l.startswith('{')][0])
m=d['ms_by_layer']; print('dbg', sys.argv[1], 'ms', round(d['ms_per_step'],2), {k: round(m[k],3) for k in ['8','9','15','21']}, d['clocks']['sm_mhz'])
PY
done
