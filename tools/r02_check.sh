set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02u_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02u_bench.json') if l.startswith('{')][0])
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
print(d['live_shape']['stack_ms'], d['live_shape']['extractor_ms'], d['live_shape']['psmnet_ms'])
PY
