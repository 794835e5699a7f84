#!/bin/bash
# every GPU command runs under its own timeout so that a hung kernel cannot eat the budget
mkdir -p gpurun_out
timeout 120 python scripts/profile_run.py cfg2 2 0 1 || { echo "cfg2 smoke failed/hung rc=$?"; exit 1; }
timeout 120 python scripts/profile_run.py cfg3 2 0 2 || { echo "cfg3 failed/hung rc=$?"; exit 1; }
timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python scripts/profile_run.py cfg3 7 0 2
