#!/bin/bash
# Runs the tc_debug experiments one process each, with a timeout, into gpurun_out/tc_debug.jsonl
out=gpurun_out/tc_debug.jsonl
: > $out
run() { timeout 90 "$@" >> $out 2>> gpurun_out/tc_debug.err || echo "{\"exp\": \"$*\", \"failed_rc\": $?}" >> $out; }
run python tools/tc_debug.py center_identity 32 32 4 16 8
run python tools/tc_debug.py tap_2_2_2 32 32 4 16 8
run python tools/tc_debug.py tap_0_1_1 32 32 4 16 8
run python tools/tc_debug.py random 32 32 4 16 8
run python tools/tc_debug.py random 64 64 5 33 17
run python tools/tc_debug.py random 32 32 40 16 8
run python tools/tc_debug.py random_epi 32 32 20 48 40
run python tools/tc_debug.py to1 32 32 6 16 8
run python tools/tc_debug.py to1 32 32 20 40 24
run python tools/tc_debug.py s2 32 64 4 32 16
run python tools/tc_debug.py s2 32 64 8 32 16
run python tools/tc_debug.py s2 64 64 40 36 20
run python tools/tc_debug.py s2 32 32 12 64 48
run python tools/tc_debug.py dec 64 32 2 16 8
run python tools/tc_debug.py dec 64 64 3 16 8
run python tools/tc_debug.py dec 64 32 7 20 12
run python tools/tc_debug.py dec 64 64 5 33 17
