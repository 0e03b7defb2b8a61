#!/bin/bash
# compute-sanitizer memcheck over every kernel family at small shapes (race / memory checking).
# usage (GPU box): bash tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck]   -> gpurun_out/sanitize_<tool>.log
tool=${1:-memcheck}
mkdir -p gpurun_out
timeout 1500 /usr/local/cuda/bin/compute-sanitizer --tool "$tool" --error-exitcode 3 --print-limit 20 python tools/sanitize_step.py > gpurun_out/sanitize_$tool.log 2>&1
rc=$?
tail -12 gpurun_out/sanitize_$tool.log
echo "compute-sanitizer $tool rc=$rc"
exit $rc
