set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02e_pytest.log
export IDISP_BENCH_SKIP_REFGPU=1
export IDISP_BENCH_SKIP_LIVE=1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
for d in 256 1; do
  IDISP_TC_DBG=$d timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_dbg$d.json 2> gpurun_out/r02e_dbg$d.err
done
python tools/show_bench.py gpurun_out/r02e_bench.json gpurun_out/r02e_dbg256.json gpurun_out/r02e_dbg1.json
timeout 300 ncu --nvtx --nvtx-include "profiled/" -k regex:conv3d_tc_kernel --launch-skip 14 --launch-count 1 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none python tools/ncu_step.py stack 32 2>&1 | grep -E "dram__|gpu__time|hit_rate"
