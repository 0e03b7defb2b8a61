#!/bin/bash
# Runs the tc_debug experiments one process each, with a timeout, into gpurun_out/tc_debug.jsonl
out=gpurun_out/tc_debug.jsonl
: > $out
run() { timeout 90 env "$@" >> $out 2>> gpurun_out/tc_debug.err || echo "{\"exp\": \"$*\", \"failed_rc\": $?}" >> $out; }
for ns in 1 0; do
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py center_identity 32 32 4 16 8
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py center_onechan 32 32 4 16 8
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py tap_1_1_0 32 32 4 16 8
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py tap_1_0_1 32 32 4 16 8
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py tap_0_1_1 32 32 4 16 8
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py tap_2_2_2 32 32 4 16 8
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py random 32 32 4 16 8
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py random 64 32 6 20 12
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py random 64 64 5 33 17
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py random 32 32 40 16 8
  run IDISP_TC_NOSTACK=$ns python tools/tc_debug.py random_epi 32 32 20 48 40
done
