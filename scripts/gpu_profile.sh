#!/bin/bash
# One ncu --set full capture of the aggregation kernels on one cfg3 frame (r=2), after the same command
# has exited 0 without ncu.  (One profiler invocation per GPU call.)
mkdir -p gpurun_out
timeout 120 python scripts/profile_run.py cfg3 2 0 1 > gpurun_out/profile_plain.json 2>gpurun_out/profile_plain.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_(hagg_split|vagg_v2|hagg_v2)' -c 4 -o gpurun_out/prof_agg python scripts/profile_run.py cfg3 2 0 1 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log; cat gpurun_out/profile_plain.json
