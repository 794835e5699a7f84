#!/bin/bash
# compute-sanitizer on the small calls of scripts/sanitize_run.py: ONE tool per GPU call (TOOL=memcheck|racecheck|synccheck),
# after the same command has exited 0 without the tool.
mkdir -p gpurun_out
TOOL=${TOOL:-memcheck}
timeout 300 python scripts/sanitize_run.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 3 python scripts/sanitize_run.py > gpurun_out/sanitize_$TOOL.log 2>&1
echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" gpurun_out/sanitize_$TOOL.log | tail -12
