set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_full.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest_full.log
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
timeout 300 ncu --nvtx --nvtx-include "profiled/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_stack.csv python tools/ncu_step.py stack 32 > /dev/null 2>&1
timeout 300 ncu --nvtx --nvtx-include "profiled/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_live.csv python tools/ncu_step.py live 8 > /dev/null 2>&1
timeout 900 ncu --nvtx --nvtx-include "profiled/" --set full --clock-control none --import-source on -f -o /tmp/r02_stack_full python tools/ncu_step.py stack 32 > /dev/null 2>&1
ncu -i /tmp/r02_stack_full.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_stack_raw.csv.gz
python tools/ncu_summary.py /tmp/r02_stack_full.ncu-rep > gpurun_out/r02_ncu_stack_summary.md
ls -la gpurun_out | tail
