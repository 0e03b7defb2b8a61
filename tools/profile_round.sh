# One round-end capture on a GPU box (gpurun): GPU tests, smoke, the driver-style bench line, launch lists (stack + live) and an
# ncu --set full capture of one configs[1] forward, everything written to gpurun_out/r02_final_* (copy to profiles/ afterwards).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
timeout 300 ncu --nvtx --nvtx-include "profiled/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_stack.csv python tools/ncu_step.py stack 32 > /dev/null 2>&1
timeout 300 ncu --nvtx --nvtx-include "profiled/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_live.csv python tools/ncu_step.py live 8 > /dev/null 2>&1
timeout 900 ncu --nvtx --nvtx-include "profiled/" --set full --clock-control none --import-source on -f -o /tmp/r02_final_stack python tools/ncu_step.py stack 32 > /dev/null 2>&1
ncu -i /tmp/r02_final_stack.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_final_ncu_stack_raw.csv.gz
python tools/ncu_summary.py /tmp/r02_final_stack.ncu-rep > gpurun_out/r02_final_ncu_stack_summary.md
ls -la gpurun_out | tail -8
