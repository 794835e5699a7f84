"""Runs the full-size frame repeatedly with ONE iteration (a single vertical + horizontal pass) and reports where the kept volume
differs between runs: localises races in the TMA / mbarrier pipelines.  usage: race_probe.py [iterations=1] [runs=6]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from stereo_matchin_b200 import api, synth
from test_gpu_parity import run_fused, P
it = int(sys.argv[1]) if len(sys.argv) > 1 else 1
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
L, R, _, D = synth.make_config("cfg3_1800x1500_d256")
ctx = api.AswContext(0)
ref = run_fused(ctx, L, R, P(ndisp=D, iterations=it), keep=True)["cost"]
for r in range(runs):
    c = run_fused(ctx, L, R, P(ndisp=D, iterations=it), keep=True)["cost"]
    bad = np.argwhere(c.view(np.uint32) != ref.view(np.uint32))
    print("run", r, "differing elements", len(bad))
    if len(bad):
        d, y, x = bad[:, 0], bad[:, 1], bad[:, 2]
        print("  d range", d.min(), d.max(), " y range", y.min(), y.max(), " x range", x.min(), x.max())
        print("  y%8", collections.Counter((y % 8).tolist()).most_common(8))
        print("  x%32", sorted(collections.Counter((x % 32).tolist()).items()))
        print("  (d-x%4)//64 task", collections.Counter((((d - x % 4)) // 64).tolist()).most_common(8))
        print("  distinct (y//8, x//32) tiles", len(set(zip((y // 8).tolist(), (x // 32).tolist()))), " rows", sorted(set(y.tolist()))[:20])
        print("  first", bad[:5].tolist())
