set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02k_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
export IDISP_BENCH_SKIP_LIVE=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"
IDISP_NO_SIDE_COPY=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02k_nosidecopy.json 2> gpurun_out/r02k_nosidecopy.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02k_bench2.json 2> gpurun_out/r02k_bench2.err
python tools/show_bench.py gpurun_out/r02k_bench.json gpurun_out/r02k_nosidecopy.json gpurun_out/r02k_bench2.json
