set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02x_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02x_bench.json') if l.startswith('{')][0])
e=d['e2e']; print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(e['value'],1), round(e['ms_per_step'],2), 'sync ms', round(e['one_batch_at_a_time']['ms_per_step'],2), d['clocks']['sm_mhz'])
print('   ', d['live_shape']['stack_ms'], d['live_shape']['extractor_ms'], d['live_shape']['psmnet_ms'])
PY
