import ctypes as C, os, sys, json
sys.path.insert(0, os.getcwd())
os.environ["ASW_B200_LIB"] = os.path.join(os.getcwd(), "stereo_matchin_b200", "libasw_b200_" + sys.argv[1] + ".so")
import numpy as np
from stereo_matchin_b200 import api, synth
L, R, _, D = synth.make_config("cfg3_1800x1500_d256")
H, W, _ = L.shape
ctx = api.AswContext(0)
dl, dr = ctx.to_device(L), ctx.to_device(R)
od = ctx.alloc(W * H)
p = api.AswParams(ndisp=D, iterations=3)
lib = api.load_library()
out = (C.c_ulonglong * 16)()
tm = ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None, timing=True)
lib.asw_debug_vprof(out)
tm = ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None, timing=True)
lib.asw_debug_vprof(out)
v = [int(x) for x in out]
n = v[8]
print("warp-steps", n, "vagg_ms", tm["vagg_mean_ms"])
names = ["before the barrier", "full-barrier wait", "step qs=0 (+finalize rows 4-7)", "step qs=1", "steps qs=2..7", "step qs=8", "step qs=9 (+finalize rows 0-3)", "finalize at tile end"]
tot = sum(v[:8])
for nm, x in zip(names, v[:8]):
    print("%-32s %8.1f cycles per warp-step (all steps)  %5.1f%%" % (nm, x / n, 100.0 * x / tot))
print("total per warp-step", tot / n)
print("per occurrence: qs0 %.0f  qs1 %.0f  qs2-7 %.0f  qs8 %.0f  qs9 %.0f cycles" % tuple(v[i] / (n / 10 * k) for i, k in ((2, 1), (3, 1), (4, 6), (5, 1), (6, 1))))

n = v[14]
if n:
    print("H pass (k_hagg_split), warp-steps", n, "hagg_ms", tm["hagg_mean_ms"])
    tot = sum(v[9:14])
    for nm, x in zip(["between steps (denominator loads issued)", "full-barrier wait", "window fill + 33 taps", "__syncthreads + issue of step m+2", "epilogue (division, stores / WTA)"], v[9:14]):
        print("%-42s %8.1f cycles per warp-step  %5.1f%%" % (nm, x / n, 100.0 * x / tot))
    print("total per warp-step", tot / n)
