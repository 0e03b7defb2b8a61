set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02s_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
export IDISP_BENCH_SKIP_LIVE=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
IDISP_NO_SIDE_COPY=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02s_nosidecopy.json 2> gpurun_out/r02s_nosidecopy.err
python tools/show_bench.py gpurun_out/r02s_bench.json gpurun_out/r02s_nosidecopy.json
