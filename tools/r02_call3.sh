set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02c_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
export IDISP_BENCH_SKIP_LIVE=1
IDISP_HEAD_OCC1=1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_headocc1.json 2> gpurun_out/r02c_headocc1.err
for d in 1 4 256 512; do
  IDISP_TC_DBG=$d timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_dbg$d.json 2> gpurun_out/r02c_dbg$d.err
done
python tools/show_bench.py gpurun_out/r02c_bench.json gpurun_out/r02c_headocc1.json gpurun_out/r02c_dbg1.json gpurun_out/r02c_dbg4.json gpurun_out/r02c_dbg256.json gpurun_out/r02c_dbg512.json
