"""One process, N GPUs: the cfg4 frame (3840x2160x256, r=7) through asw_multi_* (row bands, peer copies under the interior rows).
usage: multi_probe.py N [N ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from stereo_matchin_b200 import api, synth
L, R, _, D = synth.make_config("cfg4_3840x2160_d256")
p = api.AswParams(ndisp=D)
ref = None
for n in [int(a) for a in sys.argv[1:]]:
    with api.AswMulti(list(range(n))) as m:
        m.disparity(L, R, p, want_conf=False)
        ts = [m.disparity(L, R, p, want_conf=False) for _ in range(3)]
    t = sorted(x["timing"]["compute_ms"] for x in ts)[1]
    d = ts[0]["disp_d"]
    if ref is None:
        ref = d
    print(json.dumps({"devices": n, "compute_ms_median": t, "slowest_band_device_ms": ts[1]["timing"]["slowest_band_device_ms"],
                      "upload_ms": ts[1]["timing"]["upload_ms"], "download_ms": ts[1]["timing"]["download_ms"],
                      "Mpix_disp_per_s": L.shape[0] * L.shape[1] * D / t / 1e3, "equals_first": bool(np.array_equal(d, ref))}), flush=True)
