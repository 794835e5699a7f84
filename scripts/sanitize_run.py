"""Small hot-path / tail / cross-based calls for compute-sanitizer (memcheck): no oracle, just exercise the kernels."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_matchin_b200 import api, synth

ctx = api.AswContext(0)
for (W, H, D, it) in [(70, 40, 61, 2), (200, 60, 128, 2), (33, 9, 5, 1), (130, 35, 256, 1), (97, 50, 130, 2)]:
    L, R = synth.make_pair(W, H, D, seed=W)[:2]
    for fam in (0, 1):
        ctx.set_kernel_family(fam)
        out = ctx.disparity(L, R, api.AswParams(ndisp=D, iterations=it))
    ctx.set_kernel_family(0)
    dl, dr = ctx.to_device(L), ctx.to_device(R)
    o = ctx.alloc(W * 20)
    if H >= 30:
        ctx.disparity_raw(dl.ptr, dr.ptr, W, H, api.AswParams(ndisp=D, iterations=it), None, o.ptr, None, band=(5, 25))
        ctx.sync()
    print("hot path ok", W, H, D, flush=True)
# per-operator entry points on the TMA-fed kernels (layout conversion at the boundary)
W, H, D = 70, 40, 61
L, R = synth.make_pair(W, H, D, seed=4)[:2]
p = api.AswParams(ndisp=D)
dl, dr = ctx.to_device(L), ctx.to_device(R)
n = W * H
cost, vout, vden, hout = (ctx.alloc(4 * n * D) for _ in range(4))
tabs = [ctx.alloc(4 * n * 33) for _ in range(4)]
ctx.asw_Aggr(dl.ptr, dr.ptr, W, H, p, cost.ptr)
ctx.asw_vSupport(dl.ptr, W, H, p, tabs[0].ptr); ctx.asw_hSupport(dl.ptr, W, H, p, tabs[1].ptr)
ctx.asw_vSupport(dr.ptr, W, H, p, tabs[2].ptr); ctx.asw_hSupport(dr.ptr, W, H, p, tabs[3].ptr)
ctx.asw_vCostAggregation(W, H, p, tabs[0].ptr, tabs[2].ptr, cost.ptr, vden.ptr, vout.ptr)
ctx.asw_hCostAggregation(W, H, p, tabs[1].ptr, tabs[3].ptr, vout.ptr, None, hout.ptr)
o4 = [ctx.alloc(4 * n) for _ in range(6)]
ctx.asw_WTA(W, H, p, hout.ptr, o4[0].ptr, o4[1].ptr, o4[2].ptr, o4[3].ptr, o4[4].ptr, o4[5].ptr)
ctx.sync()
print("operators ok", flush=True)
# row bands with halo exchange: three bands on device 0 (asw_multi_*: peer copies ordered by events, two streams per band)
Lm, Rm = synth.make_pair(130, 100, 61, seed=6)[:2]
with api.AswMulti([0, 0, 0]) as m:
    om = m.disparity(Lm, Rm, api.AswParams(ndisp=61, iterations=3))
assert np.array_equal(om["disp_d"], ctx.disparity(Lm, Rm, api.AswParams(ndisp=61, iterations=3))["disp_d"])
print("multi ok", flush=True)
L, R = synth.make_pair(120, 47, 61, seed=9)[:2]
r = ctx.stereo(L, R, api.AswParams(ndisp=61, iterations=2), refine_iters=2)
c = ctx.cross_stereo(L, R)
print("whole + cross ok", r["disparity"].shape, c["final"].shape)
ctx.close()
