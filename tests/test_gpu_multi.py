"""One frame on several GPUs: row bands with a per-iteration halo exchange (asw_disparity_band_exchange_device,
asw_multi_*, stereo_matchin_b200.sharding.disparity_row_exchange_cuda) against the one-GPU frame, bit for bit.

The band threads of asw_multi only wait for each other on the HOST (no kernel spins on another kernel), so the
multi-band code path is also exercised with every band on device 0 -- that is what runs on a one-GPU box; with two or
more devices the same tests run across real devices (NVLink peer copies) and through torchrun + NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_pair
from test_gpu_parity import OP, P, assert_bit_equal, run_fused

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], [0, 0, 0, 0, 0]])
def test_multi_bands_on_one_device_equal_single(ctx, devices):
    """teddy (450 x 375, D = 61, r = 7) split into 2 / 3 / 5 bands that all live on device 0."""
    from stereo_matchin_b200.api import AswMulti
    L, R = load_pair("teddy")
    p = P()
    one = ctx.disparity(L, R, p)
    with AswMulti(devices) as m:
        out = m.disparity(L, R, p)
        again = m.disparity(L, R, p)                      # buffers and contexts are reused
    for o in (out, again):
        assert_bit_equal(o["disp_rgba"], one["disp_rgba"], "multi: disparity image")
        assert_bit_equal(o["disp_d"], one["disp_d"], "multi: disparity index")
        assert_bit_equal(o["conf"], one["conf"], "multi: confidence")
    assert out["timing"]["devices"] == len(devices) and out["timing"]["compute_ms"] > 0


def test_multi_uneven_bands_and_wide_disparity(ctx):
    """Uneven bands (H = 100 over 3), Dp > D padding (D = 130) and an odd width."""
    from stereo_matchin_b200 import synth
    from stereo_matchin_b200.api import AswMulti
    L, R = synth.make_pair(203, 100, 130, seed=7)[:2]
    p = P(ndisp=130, iterations=4)
    one = ctx.disparity(L, R, p)
    with AswMulti([0, 0, 0]) as m:
        out = m.disparity(L, R, p)
    assert_bit_equal(out["disp_d"], one["disp_d"], "multi: disparity index")
    assert_bit_equal(out["conf"], one["conf"], "multi: confidence")


def test_multi_rejects_thin_bands(ctx):
    from stereo_matchin_b200 import synth
    from stereo_matchin_b200.api import AswError, AswMulti
    L, R = synth.make_pair(64, 40, 16, seed=3)[:2]
    with AswMulti([0, 0, 0]) as m:                       # 13-row bands < radius
        with pytest.raises(AswError):
            m.disparity(L, R, P(ndisp=16, iterations=2))


def test_band_exchange_callback_contract(ctx):
    """Python-level callback: a band without neighbours gets NULL pointers; a band in the middle of the frame gets four
    device pointers `radius` volume rows long.  Emulating both neighbours with the one-GPU volume is not possible from
    here (they would have to run in lock step), so this test only checks the contract; the data path is covered above."""
    from stereo_matchin_b200 import synth
    L, R = synth.make_pair(96, 80, 61, seed=5)[:2]
    H, W, _ = L.shape
    p = P(iterations=3)
    dl, dr = ctx.to_device(L), ctx.to_device(R)
    out = ctx.alloc(W * H)
    calls = []
    ctx.disparity_band_exchange(dl.ptr, dr.ptr, W, H, p, (0, H), None, out.ptr, None, lambda *a: calls.append(a))
    ctx.sync()
    assert [c[0] for c in calls] == [0, 1] and all(c[1:5] == (None, None, None, None) for c in calls)
    full = run_fused(ctx, L, R, p)
    assert_bit_equal(out.download((H, W), np.uint8), full["d"], "exchange entry, whole frame")
    calls.clear()
    ctx.disparity_band_exchange(dl.ptr, dr.ptr, W, H, p, (32, 56), None, out.ptr, None, lambda *a: calls.append(a))
    ctx.sync()
    it, ts, bs, tr, br, nbytes = calls[0]
    Wv, Dp = ((W + 63) // 64) * 64 + 32, 64
    assert nbytes == 16 * Wv * Dp * 4 and None not in (ts, bs, tr, br)
    assert ts - tr == nbytes and br - bs == nbytes and bs - ts == (24 - 16) * Wv * Dp * 4


@pytest.mark.skipif("_ndev() < 2")
def test_multi_real_devices_equal_single(ctx):
    from stereo_matchin_b200 import synth
    from stereo_matchin_b200.api import AswMulti
    L, R = synth.make_pair(640, 360, 128, seed=11)[:2]
    p = P(ndisp=128)
    one = ctx.disparity(L, R, p)
    for devs in ([0, 1], list(range(min(_ndev(), 8)))):
        with AswMulti(devs) as m:
            out = m.disparity(L, R, p)
        assert_bit_equal(out["disp_d"], one["disp_d"], f"devices {devs}: disparity index")
        assert_bit_equal(out["conf"], one["conf"], f"devices {devs}: confidence")


@pytest.mark.skipif("_ndev() < 2")
def test_nccl_row_exchange_equals_single():
    """torchrun, one process per GPU, NCCL send/recv of the halo rows + all-gather of the bands (tests/mgpu_worker.py)."""
    n = min(_ndev(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29653", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "equals_1gpu True" in r.stdout


# ---------------------------------------------------------------------------------------------
# BASELINE.json shapes that round 1 did not cover: cfg4 width (3840, D = 256) and cfg5 (1280 x 720, D = 128)

def test_cfg4_width_band_vs_oracle(ctx, oracle):
    """16 rows in the middle of the 3840 x 2160 x 256 frame (120 x-blocks, 4 disparity tasks) against the CPU oracle run
    on the rows that can influence them (r * R = 112 halo rows each side)."""
    from stereo_matchin_b200.synth import make_config
    L, R, _, D = make_config("cfg4_3840x2160_d256")
    p = P(ndisp=D)
    y0, y1 = 1000, 1016
    b = run_fused(ctx, L, R, p, band=(y0, y1))
    ya, yb = y0 - 112, y1 + 112
    o = oracle.asw_hot_path(np.ascontiguousarray(L[ya:yb]), np.ascontiguousarray(R[ya:yb]), OP(p), use_fma=True)
    assert_bit_equal(b["d"].astype(np.float32), o["d_ref"][112:128], "cfg4 band vs oracle: disparity")
    assert_bit_equal(b["conf"], o["conf_ref"][112:128], "cfg4 band vs oracle: confidence")


def test_cfg5_frame_vs_oracle(ctx, oracle):
    """A whole 1280 x 720 x 128 frame of the cfg5 batch (Dp = 128: one 128-disparity window) against the CPU oracle."""
    from stereo_matchin_b200.synth import make_config
    L, R, _, D = make_config("cfg5_1280x720_d128", index=3)
    p = P(ndisp=D)
    g = run_fused(ctx, L, R, p)
    o = oracle.asw_hot_path(L, R, OP(p), use_fma=True)
    assert_bit_equal(g["d"].astype(np.float32), o["d_ref"], "cfg5 frame vs oracle: disparity")
    assert_bit_equal(g["conf"], o["conf_ref"], "cfg5 frame vs oracle: confidence")
    assert_bit_equal(g["left"], o["left"], "cfg5 frame vs oracle: disparity image")


def test_host_binary_devices_option(tmp_path):
    """src/host/stereo_matching --devices 0,0 --method hot: every frame is split over the listed devices (here: two bands
    on device 0) and the PNG equals the one written with --device 0."""
    import shutil
    from conftest import GOLDEN, PAIRS, load_rgba
    exe = os.path.join(ROOT, "src", "host", "stereo_matching")
    assert os.path.exists(exe), "run `python -m stereo_matchin_b200.build` first"
    os.makedirs(tmp_path / "teddy")
    for f in PAIRS["teddy"]:
        shutil.copy(os.path.join(GOLDEN, "teddy", f), tmp_path / "teddy" / f)
    (tmp_path / "pics.txt").write_text("teddy/im2.png\nteddy/im6.png\n")
    outs = {}
    for tag, dev in (("one", ["--device", "0"]), ("two", ["--devices", "0,0"])):
        r = subprocess.run([exe, "--pics", str(tmp_path / "pics.txt"), "--root", str(tmp_path), "--runs", "2", "--method", "hot",
                            "--out-suffix", "_" + tag, "--log", str(tmp_path / (tag + ".tsv"))] + dev, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[tag] = load_rgba(str(tmp_path / "teddy" / ("asw_disparity_" + tag + ".png")))
    assert np.array_equal(outs["one"], outs["two"])
