"""Small hot-path / tail / cross-based calls for compute-sanitizer (memcheck): no oracle, just exercise the kernels."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_matchin_b200 import api, synth

ctx = api.AswContext(0)
for (W, H, D, it) in [(70, 40, 61, 2), (200, 60, 128, 2), (33, 9, 5, 1), (130, 35, 256, 1), (97, 50, 130, 2)]:
    L, R = synth.make_pair(W, H, D, seed=W)[:2]
    for fam in (0, 2, 1):
        ctx.set_kernel_family(fam)
        out = ctx.disparity(L, R, api.AswParams(ndisp=D, iterations=it))
    ctx.set_kernel_family(0)
    dl, dr = ctx.to_device(L), ctx.to_device(R)
    o = ctx.alloc(W * 20)
    if H >= 30:
        ctx.disparity_raw(dl.ptr, dr.ptr, W, H, api.AswParams(ndisp=D, iterations=it), None, o.ptr, None, band=(5, 25))
        ctx.sync()
    print("hot path ok", W, H, D, flush=True)
L, R = synth.make_pair(120, 47, 61, seed=9)[:2]
r = ctx.stereo(L, R, api.AswParams(ndisp=61, iterations=2), refine_iters=2)
c = ctx.cross_stereo(L, R)
print("whole + cross ok", r["disparity"].shape, c["final"].shape)
ctx.close()
