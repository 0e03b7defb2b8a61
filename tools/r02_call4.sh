set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02d_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
export IDISP_BENCH_SKIP_LIVE=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
for d in 1024 64 128; do
  IDISP_TC_DBG=$d timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_dbg$d.json 2> gpurun_out/r02d_dbg$d.err
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02d_bench2.json 2> gpurun_out/r02d_bench2.err
python tools/show_bench.py gpurun_out/r02d_bench.json gpurun_out/r02d_bench2.json gpurun_out/r02d_dbg1024.json gpurun_out/r02d_dbg64.json gpurun_out/r02d_dbg128.json
