#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for ny in 8 16; do ASW_V_NY=$ny python scripts/profile_run.py cfg3 2 0 2; done
python scripts/profile_run.py cfg3 7 0 2
