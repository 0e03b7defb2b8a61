set -x
mkdir -p gpurun_out
timeout 600 ncu --nvtx --nvtx-include "profiled/" -k regex:conv3d_tc_kernel --launch-skip 5 --launch-count 1 --set full --clock-control none --import-source on -f -o /tmp/r02_s2t python tools/ncu_step.py stack 32 > /dev/null 2>&1
ncu -i /tmp/r02_s2t.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_s2t_source.csv.gz
ncu -i /tmp/r02_s2t.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_s2t_raw.csv.gz
ls -la gpurun_out/ | tail -3
