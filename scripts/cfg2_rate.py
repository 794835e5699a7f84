"""Back-to-back calls of the hot path on a small frame (default cfg2, the reference's 450x375x61 shape): ms per frame with the
CUDA-graph replay (default) or kernel-by-kernel launches (ASW_GRAPH=0).  usage: cfg2_rate.py [workload=cfg2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_matchin_b200 import api, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
name = {"cfg2": "cfg2_teddy_shape", "cfg5": "cfg5_1280x720_d128"}[wl]
L, R, _, D = synth.make_config(name)
H, W, _ = L.shape
ctx = api.AswContext(0)
dl, dr = ctx.to_device(L), ctx.to_device(R)
od = ctx.alloc(W * H)
p = api.AswParams(ndisp=D, iterations=7)
for _ in range(5):
    ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None)
ctx.sync()
best = 1e9
for rep in range(5):
    t0 = time.perf_counter()
    for _ in range(50):
        ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None)
    ctx.sync()
    best = min(best, (time.perf_counter() - t0) / 50 * 1e3)
print("%s graph=%s: %.4f ms/frame  %.0f Mpix*disp/s" % (wl, os.environ.get("ASW_GRAPH", "1"), best, W * H * D / best / 1e3))
