set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02f_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02f_bench.json') if l.startswith('{')][0])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
print('live', d.get('live_shape'))
PY
