import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from stereo_matchin_b200 import api, synth
from test_gpu_parity import run_fused
ctx = api.AswContext(0)
L, R, _, D = synth.make_config("cfg3_1800x1500_d256")
for it in (1, 2, 7):
    p = api.AswParams(ndisp=D, iterations=it)
    f1 = run_fused(ctx, L, R, p); f2 = run_fused(ctx, L, R, p)
    print("it", it, "full vs full: d diff", (f1["d"] != f2["d"]).sum(), "conf diff", (f1["conf"].view(np.uint32) != f2["conf"].view(np.uint32)).sum())
    for band in ((700, 716), (696, 720), (0, 16), (1484, 1500)):
        b = run_fused(ctx, L, R, p, band=band)
        dd = b["conf"].view(np.uint32) != f1["conf"][band[0]:band[1]].view(np.uint32)
        ys, xs = np.nonzero(dd)
        print("  band", band, "conf diff", dd.sum(), "rows", np.unique(ys)[:10], "x range", (xs.min(), xs.max()) if len(xs) else None)
    fb = run_fused(ctx, L, R, p, family=1, band=(700, 716))
    for name, g in (("full", f1["conf"][700:716]), ):
        print("  basic-band vs", name, (fb["conf"].view(np.uint32) != g.view(np.uint32)).sum())
