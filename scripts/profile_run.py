"""One hot-path invocation for profiling (ncu) and per-stage timing.
usage: profile_run.py [workload=cfg3] [iterations=2] [family=0] [reps=1]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from stereo_matchin_b200 import api, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
it = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fam = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
name = {"cfg2": "cfg2_teddy_shape", "cfg3": "cfg3_1800x1500_d256", "cfg4": "cfg4_3840x2160_d256", "cfg5": "cfg5_1280x720_d128"}[wl]
L, R, _, D = synth.make_config(name)
H, W, _ = L.shape
ctx = api.AswContext(0)
ctx.set_kernel_family(fam)
dl, dr = ctx.to_device(L), ctx.to_device(R)
od = ctx.alloc(W * H)
p = api.AswParams(ndisp=D, iterations=it)
for _ in range(reps):
    tm = ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None, timing=True)
tm["Mpix_disp_per_s"] = W * H * D / tm["total_ms"] / 1e3
tm["workload"] = f"{wl} {W}x{H}x{D} r={it} family={fam}"
print(json.dumps(tm))
