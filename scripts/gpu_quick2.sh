#!/bin/bash
mkdir -p gpurun_out
timeout 60 python scripts/profile_run.py cfg2 2 0 1 || { echo "cfg2 failed/hung rc=$?"; exit 1; }
timeout 60 python -c "
import numpy as np, sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from stereo_matchin_b200 import api, synth
from oracle import asw_oracle as o
ctx = api.AswContext(0)
L, R = synth.make_pair(200, 60, 128, seed=3)[:2]
p = api.AswParams(ndisp=128, iterations=2)
out = ctx.disparity(L, R, p)
ref = o.asw_hot_path(L, R, o.OracleParams(16, 128, 30.91, 28.21, float('inf'), 2), use_fma=True)
print('small TMA case equal:', np.array_equal(out['disp_rgba'], ref['left']), np.array_equal(out['conf'].view(np.uint32), ref['conf_ref'].view(np.uint32)))
" || { echo "small parity failed/hung rc=$?"; exit 1; }
timeout 120 python scripts/profile_run.py cfg3 2 0 2 || { echo "cfg3 failed/hung rc=$?"; exit 1; }
timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python scripts/profile_run.py cfg3 7 0 2
