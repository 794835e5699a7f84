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
out = (C.c_ulonglong * 8)()
tm = ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None, timing=True)
lib.asw_debug_vprof(out)
tm = ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None, timing=True)
lib.asw_debug_vprof(out)
v = [int(x) for x in out]
n = v[4]
print("warp-steps", n, "vagg_ms", tm["vagg_mean_ms"])
names = ["bookkeeping(before wait)", "full-barrier wait", "cost loads + math + arrive", "epilogue", ]
tot = sum(v[:4])
for nm, x in zip(names, v[:4]):
    print("%-30s %8.1f cycles/warp-step  %5.1f%%" % (nm, x / n, 100.0 * x / tot))
print("total per warp-step", tot / n)
