#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cross.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_cross.log 2>&1; echo "cross rc=$?"; tail -25 gpurun_out/pytest_cross.log
