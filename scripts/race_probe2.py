"""Instrumented build only: are the vertical pass' stored denominators (written by iteration 0) identical between runs?"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASW_B200_LIB"] = os.path.join(os.getcwd(), "stereo_matchin_b200", "libasw_b200_vprof.so")
import numpy as np
from stereo_matchin_b200 import api, synth
L, R, _, D = synth.make_config("cfg3_1800x1500_d256")
H, W, _ = L.shape
ctx = api.AswContext(0)
lib = api.load_library()
lib.asw_debug_vden.restype = C.c_longlong
lib.asw_debug_vden.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
dl, dr = ctx.to_device(L), ctx.to_device(R)
od = ctx.alloc(W * H)
p = api.AswParams(ndisp=D, iterations=2)
ref = None
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None)
    n = lib.asw_debug_vden(ctx.h, None, 0)
    n = lib.asw_debug_vden(ctx.h, None, 1 << 40)
    buf = np.empty(n, np.float32)
    lib.asw_debug_vden(ctx.h, buf.ctypes.data, n)
    if ref is None: ref = buf
    else:
        bad = np.flatnonzero(buf.view(np.uint32) != ref.view(np.uint32))
        print("run", r, "den floats", n, "differing", len(bad), bad[:8].tolist(), [(int(b) // 4) % 256 for b in bad[:4]], [((int(b) // 4) // 256) % 64 for b in bad[:4]])
