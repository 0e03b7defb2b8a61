set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02b_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
export IDISP_BENCH_SKIP_LIVE=1
for d in 1 4 256; do
  IDISP_TC_DBG=$d timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_dbg$d.json 2> gpurun_out/r02b_dbg$d.err
done
python tools/show_bench.py gpurun_out/r02b_bench.json gpurun_out/r02b_dbg1.json gpurun_out/r02b_dbg4.json gpurun_out/r02b_dbg256.json
