set -x
mkdir -p gpurun_out
# source-level captures: layer 1 (32->32 per-step-triple kernel) and conv6 (transposed, DTR), configs[1] B=32
timeout 600 ncu --nvtx --nvtx-include "profiled/" -k regex:conv3d_tc_kernel --launch-skip 2 --launch-count 1 --set full --clock-control none --import-source on -f -o /tmp/r02_tri python tools/ncu_step.py stack 32 > /dev/null 2>&1
ncu -i /tmp/r02_tri.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_tri_source.csv.gz
ncu -i /tmp/r02_tri.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_tri_raw.csv.gz
timeout 600 ncu --nvtx --nvtx-include "profiled/" -k regex:conv3d_tc_kernel --launch-skip 14 --launch-count 1 --set full --clock-control none --import-source on -f -o /tmp/r02_dtr python tools/ncu_step.py stack 32 > /dev/null 2>&1
ncu -i /tmp/r02_dtr.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_dtr_source.csv.gz
ncu -i /tmp/r02_dtr.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_dtr_raw.csv.gz
ls -la gpurun_out/
