set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02t_pytest.log
bash tools/sanitize.sh memcheck
export IDISP_BENCH_SKIP_REFGPU=1
export IDISP_BENCH_SKIP_LIVE=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02t_bench.json') if l.startswith('{')][0])
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roi_align']['ms_per_call'], d['clocks']['sm_mhz'])
PY
