set -x
mkdir -p gpurun_out
export IDISP_BENCH_SKIP_REFGPU=1
export IDISP_BENCH_SKIP_LIVE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02o_bench_2gpu.json 2> gpurun_out/r02o_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r02o_bench_2gpu.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02o_bench_1gpu.json 2> gpurun_out/r02o_bench_1gpu.err
python - <<'PY'
import json
for f in ['r02o_bench_2gpu','r02o_bench_1gpu']:
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][0])
        print(f, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d.get('allgather'), d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 python -m pytest tests -m gpu -q -k "two_rank" 2>&1 | tail -3
