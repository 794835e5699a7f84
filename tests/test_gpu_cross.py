"""Parity of the cross-based method (the reference's second pipeline, main.cpp:258-367) through the C ABI
against the CPU oracle (oracle/cross_oracle.c; every output compared BIT-EXACT, float volumes included)
and against the reference's committed cross_based_initial.png / cross_based_disparity.png.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, crop_pair, load_pair, load_rgba
from test_gpu_parity import assert_bit_equal

pytestmark = pytest.mark.gpu


def CP(**kw):
    from stereo_matchin_b200.api import CrossParams
    return CrossParams(**kw)


CASES = [("teddy", 0, 0, 96, 40, 61, 25), ("cones", 300, 200, 150, 37, 61, 25), ("art", 10, 10, 33, 70, 5, 4),
         ("laundry", 440, 0, 10, 50, 9, 25), ("teddy", 7, 9, 1, 35, 3, 2), ("teddy", 7, 9, 35, 1, 3, 25),
         ("tsukuba", 0, 0, 300, 20, 16, 1), ("cones", 0, 0, 70, 70, 1, 25)]


@pytest.mark.parametrize("ds,x0,y0,w,h,D,arm", CASES)
def test_cross_operator_parity(ctx, oracle, ds, x0, y0, w, h, D, arm):
    L, R = crop_pair(ds, x0, y0, w, h)
    p = CP(ndisp=D, max_arm=arm)
    n = w * h
    dl, dr = ctx.to_device(L), ctx.to_device(R)
    ml, mr = ctx.alloc(4 * n), ctx.alloc(4 * n)
    for local in (1, 3):
        ctx.asw_Median_grid(w, h, local, dl.ptr, ml.ptr)
        assert_bit_equal(ml.download((h, w, 4), np.uint8), oracle.cb_median_grid(L, local), f"Median grid local={local}")
    ctx.asw_Median_grid(w, h, 3, dr.ptr, mr.ptr)
    oml, omr = oracle.cb_median_grid(L, 3), oracle.cb_median_grid(R, 3)

    cl, cr = ctx.alloc(16 * n), ctx.alloc(16 * n)
    ctx.cb_op("asw_Cross", w, h, p, ml.ptr, cl.ptr)                                   # cross.cl
    ctx.cb_op("asw_Cross", w, h, p, mr.ptr, cr.ptr)
    ocl, ocr = oracle.cb_cross(oml, arm), oracle.cb_cross(omr, arm)
    assert_bit_equal(cl.download((4, h, w), np.int32), ocl, "Cross left")
    assert_bit_equal(cr.download((4, h, w), np.int32), ocr, "Cross right")

    cost, tmp = ctx.alloc(4 * n * D), ctx.alloc(4 * n * D)
    ctx.cb_op("asw_Aggregation", w, h, p, ml.ptr, mr.ptr, cost.ptr)                   # aggregation.cl
    oc = oracle.cb_aggregation(oml, omr, D)
    assert_bit_equal(cost.download((D, h, w), np.float32), oc, "Aggregation")
    ctx.cb_op("asw_Integral_h", w, h, p, cost.ptr)                                    # integral_h.cl
    oih = oracle.cb_integral(oc, True)
    assert_bit_equal(cost.download((D, h, w), np.float32), oih, "Integral_h")
    ctx.cb_op("asw_Oii_hcross", w, h, p, cl.ptr, cr.ptr, cost.ptr, tmp.ptr)           # oii_hcross.cl
    ooh = oracle.cb_oii(ocl, ocr, oih, True)
    assert_bit_equal(tmp.download((D, h, w), np.float32), ooh, "Oii_hcross")
    ctx.cb_op("asw_Integral_v", w, h, p, tmp.ptr)                                     # integral_v.cl
    oiv = oracle.cb_integral(ooh, False)
    assert_bit_equal(tmp.download((D, h, w), np.float32), oiv, "Integral_v")
    ctx.cb_op("asw_Oii_vcross", w, h, p, cl.ptr, cr.ptr, tmp.ptr, cost.ptr)           # oii_vcross.cl
    oov = oracle.cb_oii(ocl, ocr, oiv, False)
    assert_bit_equal(cost.download((D, h, w), np.float32), oov, "Oii_vcross")

    init, voted = ctx.alloc(4 * n), ctx.alloc(4 * n)
    ctx.cb_op("asw_Init_disparity", w, h, p, cost.ptr, init.ptr)                      # init_disparity.cl
    oinit = oracle.cb_init_disparity(oov)
    assert_bit_equal(init.download((h, w, 4), np.uint8), oinit, "Init_disparity")
    ctx.cb_op("asw_Disparity", w, h, p, init.ptr, cl.ptr, voted.ptr)                  # disparity.cl
    assert_bit_equal(voted.download((h, w, 4), np.uint8), oracle.cb_disparity(oinit, ocl, D), "Disparity")


@pytest.mark.parametrize("ds,D,local", [("tsukuba", 61, 3), ("art", 61, 3), ("sukub", 16, 1), ("laundry", 61, 3)])
def test_cross_whole_method_parity(ctx, oracle, ds, D, local):
    L, R = load_pair(ds)
    p = CP(ndisp=D, median_local=local)
    got = ctx.cross_stereo(L, R, p)
    want = oracle.cross_full(L, R, D=D, median_local=local)
    for key in ("median_l", "initial", "final"):
        assert_bit_equal(got[key], want[key], f"asw_cross_stereo {key}")
    t = got["timing"]
    assert t["total_ms"] > 0 and all(t[k] > 0 for k in ("median_ms", "cross_ms", "aggregation_ms", "integral_h_ms", "oii_h_ms",
                                                         "integral_v_ms", "oii_v_ms", "init_disparity_ms", "final_disparity_ms"))


@pytest.mark.parametrize("ds", ["teddy", "art"])
def test_cross_whole_method_vs_reference_png(ctx, ds):
    """The shipped path against the PNGs the reference committed (teddy / art: the final image is byte-exact)."""
    L, R = load_pair(ds)
    got = ctx.cross_stereo(L, R)
    assert np.array_equal(got["final"], load_rgba(os.path.join(GOLDEN, ds, "cross_based_disparity.png")))
    gi = load_rgba(os.path.join(GOLDEN, ds, "cross_based_initial.png"))
    assert int((gi != got["initial"]).any(-1).sum()) <= 20


def test_cross_error_behaviour(ctx):
    from stereo_matchin_b200.api import AswError
    L, R = crop_pair("teddy", 0, 0, 32, 16)
    with pytest.raises(AswError):
        ctx.cross_stereo(L, R, CP(ndisp=0))
    with pytest.raises(AswError):
        ctx.cross_stereo(L, R, CP(ndisp=300))
    with pytest.raises(AswError):
        ctx.cross_stereo(L, R, CP(max_arm=0))


def test_host_binary_both_methods(tmp_path):
    """The pics.txt-driven host program runs the reference's full per-run sequence (cross-based + ASW)."""
    import shutil
    import subprocess
    from conftest import PAIRS
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "src", "host", "stereo_matching")
    assert os.path.exists(exe), "run `python -m stereo_matchin_b200.build` first"
    os.makedirs(tmp_path / "teddy")
    for f in PAIRS["teddy"]:
        shutil.copy(os.path.join(GOLDEN, "teddy", f), tmp_path / "teddy" / f)
    (tmp_path / "pics.txt").write_text("teddy/im2.png\nteddy/im6.png\n")
    r = subprocess.run([exe, "--pics", str(tmp_path / "pics.txt"), "--root", str(tmp_path), "--runs", "1", "--method", "both",
                        "--out-suffix", "", "--log", str(tmp_path / "log.tsv")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(load_rgba(str(tmp_path / "teddy" / "cross_based_disparity.png")),
                          load_rgba(os.path.join(GOLDEN, "teddy", "cross_based_disparity.png")))
    for name in ("cross_based_initial.png", "median.png", "asw_disparity.png", "asw_consistency_pre-reff.png", "asw_consistency_post-reff.png"):
        assert os.path.exists(tmp_path / "teddy" / name), name
    run1 = [ln for ln in (tmp_path / "log.tsv").read_text().splitlines() if ln.startswith("Run 1")][0].split("\t")
    vals = [float(v) for v in run1[1:] if v.strip()]
    assert len(vals) == 30 and all(v > 0 for v in vals)      # every column of the reference's log is filled
