"""Times one rank's share of the 4K frame for different sharding grids on a single GPU (worst-case interior rank)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stereo_matchin_b200 import api, synth
W, H, D, _ = synth.CONFIGS["cfg4_3840x2160_d256"]
L, R = synth.make_config("cfg4_3840x2160_d256", 0)[:2]
ctx = api.AswContext(0)
p = api.AswParams(ndisp=D, iterations=7)
dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
for name, band, ds in [("8x1 interior band", (810, 1080), (0, 256)), ("4x2 interior band", (540, 1080), (0, 128)),
                       ("2x4 band", (0, 1080), (64, 128)), ("1x4 (4 ranks)", (0, 2160), (64, 128)), ("2x2 (4 ranks)", (0, 1080), (0, 128))]:
    rows = band[1] - band[0]
    m = torch.empty((3, rows, W), dtype=torch.float32, device="cuda")
    best = None
    for _ in range(3):
        t = ctx.disparity_shard_raw(dl.data_ptr(), dr.data_ptr(), W, H, p, band, ds, m[0].data_ptr(), m[1].data_ptr(), m[2].data_ptr(), timing=True)
        best = t if best is None or t["total_ms"] < best["total_ms"] else best
    print(name, json.dumps({k: round(v, 3) for k, v in best.items() if k.endswith("_ms") and v}), flush=True)
del dl, dr, m
torch.cuda.synchronize(); torch.cuda.empty_cache(); ctx.close()
