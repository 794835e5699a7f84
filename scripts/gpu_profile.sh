#!/bin/bash
# ncu launch list + full capture of the aggregation kernels on one cfg3 frame (r=2).
mkdir -p gpurun_out
echo skip-pytest
timeout 120 python scripts/profile_run.py cfg3 2 0 2 > gpurun_out/profile_plain.json 2>gpurun_out/profile_plain.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_run.py cfg3 2 0 1 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; cat gpurun_out/profile_plain.json
timeout 120 python scripts/profile_run.py cfg3 2 0 1 > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_(hagg_split|vagg_v2|hagg_v2)' -c 4 -o gpurun_out/prof_agg python scripts/profile_run.py cfg3 2 0 1 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
