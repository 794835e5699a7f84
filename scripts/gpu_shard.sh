#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/profile_run.py cfg2 2 0 1 || { echo "cfg2 failed/hung"; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 -k "shard" > gpurun_out/pytest_shard.log 2>&1; echo "shard rc=$?"; tail -15 gpurun_out/pytest_shard.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python scripts/profile_run.py cfg3 7 0 2
