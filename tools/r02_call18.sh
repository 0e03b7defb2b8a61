set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02r_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02r_again.json 2> gpurun_out/r02r_again.err
python - <<'PY'
import json
for f in ['r02r_bench','r02r_again']:
    d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][0])
    print(f, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])
    print('   live', d['live_shape']['stack_ms'], d['live_shape']['psmnet_ms'])
PY
