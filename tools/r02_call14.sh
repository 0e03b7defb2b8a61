set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02m_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
export IDISP_BENCH_SKIP_LIVE=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"
IDISP_TC_DBG=2048 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02m_nosplit.json 2> gpurun_out/r02m_nosplit.err
python tools/show_bench.py gpurun_out/r02m_bench.json gpurun_out/r02m_nosplit.json
IDISP_SOFTARGMIN_NO_CELLMAX=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02m_nocellmax.json 2> gpurun_out/r02m_nocellmax.err
python tools/show_bench.py gpurun_out/r02m_bench.json gpurun_out/r02m_nosplit.json gpurun_out/r02m_nocellmax.json
