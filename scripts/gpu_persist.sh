#!/bin/bash
mkdir -p gpurun_out
timeout 60 python scripts/profile_run.py cfg2 2 0 1 || { echo "cfg2 failed/hung rc=$?"; exit 1; }
timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in cfg3 cfg2 cfg5; do for p in 0 1; do echo -n "$wl persist=$p: "; ASW_V_PERSIST=$p timeout 120 python scripts/profile_run.py $wl 7 0 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(d[k],3) for k in ('vagg_mean_ms','hagg_mean_ms','total_ms')})"; done; done
for p in 0 1; do echo "shards persist=$p"; ASW_V_PERSIST=$p timeout 200 python scripts/shard_probe.py | cut -c1-200; done
