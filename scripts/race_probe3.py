"""Determinism stress: the full-size frame N times with 7 iterations, disparity + confidence maps compared with the first run.
usage: race_probe3.py [runs=40]   (ASW_B200_LIB selects the library build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereo_matchin_b200 import api, synth
L, R, _, D = synth.make_config("cfg3_1800x1500_d256")
H, W, _ = L.shape
ctx = api.AswContext(0)
dl, dr = ctx.to_device(L), ctx.to_device(R)
od, oc = ctx.alloc(W * H), ctx.alloc(W * H * 4)
p = api.AswParams(ndisp=D, iterations=7)
ref = None
bad_runs = 0
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, oc.ptr)
    ctx.sync()
    c = oc.download((H, W), np.float32).view(np.uint32)
    if ref is None: ref = c
    else:
        nb = int((c != ref).sum())
        if nb:
            bad_runs += 1
            ys, xs = np.nonzero(c != ref)
            print("  run", r, "conf differs at", nb, "pixels; y%8", sorted(set((ys % 8).tolist())), "x range", xs.min(), xs.max(), "y range", ys.min(), ys.max())
print("bad runs", bad_runs)
