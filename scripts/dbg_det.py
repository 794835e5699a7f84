import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from stereo_matchin_b200 import api, synth
from test_gpu_parity import run_fused
ctx = api.AswContext(0)
L, R, _, D = synth.make_config("cfg3_1800x1500_d256")
for it in (1, 2, 3, 7):
    p = api.AswParams(ndisp=D, iterations=it)
    f = [run_fused(ctx, L, R, p, keep=(it <= 2)) for _ in range(3)]
    for a in (1, 2):
        line = f"it {it} run0 vs run{a}: d diff {(f[0]['d'] != f[a]['d']).sum()} conf diff {(f[0]['conf'].view(np.uint32) != f[a]['conf'].view(np.uint32)).sum()}"
        if it <= 2:
            dc = f[0]['cost'].view(np.uint32) != f[a]['cost'].view(np.uint32)
            dd, ys, xs = np.nonzero(dc)
            line += f" cost diff {dc.sum()}"
            if len(dd): line += f" d range {dd.min()}-{dd.max()} y {ys.min()}-{ys.max()} x {xs.min()}-{xs.max()} sample {(dd[0], ys[0], xs[0])}"
        print(line)
# detail for it=2
p = api.AswParams(ndisp=D, iterations=2)
a = run_fused(ctx, L, R, p, keep=True); b = run_fused(ctx, L, R, p, keep=True); c = run_fused(ctx, L, R, p, family=1, keep=True)
for name, u in (("run a", a), ("run b", b)):
    dc = u['cost'].view(np.uint32) != c['cost'].view(np.uint32)
    dd, ys, xs = np.nonzero(dc)
    print(name, "vs basic kernels: cost diff", dc.sum())
    if len(dd):
        print("  d hist (by 32):", np.bincount(dd // 32, minlength=8))
        print("  y%8 hist:", np.bincount(ys % 8, minlength=8))
        print("  x%32 hist:", np.bincount(xs % 32, minlength=32))
        print("  first:", list(zip(dd[:5], ys[:5], xs[:5])))
