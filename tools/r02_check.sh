set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02z_bench_4gpu.json 2> gpurun_out/r02z_bench_4gpu.err; echo "bench4 rc=$?"
tail -2 gpurun_out/r02z_bench_4gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02z_bench_4gpu.json') if l.startswith('{')][0])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d.get('allgather'), d['clocks'])
PY
