"""Parity of the consumers of the hot path (consistency, refinement, penalised WTA, median) and of
the whole method (asw_stereo), called through the C ABI, against the CPU oracle
(oracle/asw_tail_oracle.c, use_fma=1: the same operations in the same order, so every output --
float planes included -- is compared BIT-EXACT) and against the reference's committed PNGs.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, crop_pair, load_pair, load_rgba
from test_gpu_parity import OP, P, assert_bit_equal

pytestmark = pytest.mark.gpu


def grey_rgba(d, D):
    """Disparity indices -> the RGBA8 image asw_WTA writes (grey = q8(d/(D-1)), A = 255)."""
    f = (d.astype(np.float32) / np.float32(max(D - 1, 1))) * np.float32(255.0)
    v = np.ceil(f - np.float32(0.5)).clip(0, 255).astype(np.uint8)      # round half down, like q8
    return np.stack([v, v, v, np.full_like(v, 255)], -1)


TAIL_CASES = [("teddy", 0, 0, 96, 40, 61, 16), ("cones", 300, 200, 150, 37, 61, 16), ("art", 10, 10, 33, 70, 5, 4),
              ("laundry", 440, 0, 10, 50, 9, 16), ("teddy", 7, 9, 1, 35, 3, 2), ("teddy", 7, 9, 35, 1, 3, 16),
              ("tsukuba", 0, 0, 300, 20, 16, 0)]


@pytest.mark.parametrize("ds,x0,y0,w,h,D,R", TAIL_CASES)
def test_tail_operator_parity(ctx, oracle, ds, x0, y0, w, h, D, R):
    L, Rt = crop_pair(ds, x0, y0, w, h)
    rng = np.random.default_rng(w * 131 + h)
    p = P(ndisp=D, radius=R)
    n = w * h
    dl, dr = ctx.to_device(L), ctx.to_device(Rt)
    d_l = rng.integers(0, D, (h, w))
    d_r = np.where(rng.random((h, w)) < 0.7, d_l, rng.integers(0, D, (h, w)))
    lw, rw = grey_rgba(d_l, D), grey_rgba(d_r, D)
    conf_l, conf_r = rng.random((h, w), dtype=np.float32), rng.random((h, w), dtype=np.float32)
    dscale = float(D - 1)

    # Constistency (consist.cl)
    b_lw, b_rw, b_cl, b_cr = ctx.to_device(lw), ctx.to_device(rw), ctx.to_device(conf_l), ctx.to_device(conf_r)
    b_ce, b_red = ctx.alloc(4 * n), ctx.alloc(4 * n)
    ctx.asw_Constistency(w, h, p, b_lw.ptr, b_rw.ptr, b_cl.ptr, b_cr.ptr, b_ce.ptr, b_red.ptr)
    ocl, ocr = conf_l.copy(), conf_r.copy()
    oce, ored = oracle.consistency(lw, rw, ocl, ocr, dscale=dscale)
    assert_bit_equal(b_ce.download((h, w, 4), np.uint8), oce, "Constistency output")
    assert_bit_equal(b_red.download((h, w, 4), np.uint8), ored, "Constistency output_red")
    assert_bit_equal(b_cl.download((h, w), np.float32), ocl, "Constistency confidence_ref")
    assert_bit_equal(b_cr.download((h, w), np.float32), ocr, "Constistency confidence_tar")

    # asw_ref_v / asw_ref_h (asw_refinement_v.cl / asw_refinement_h.cl), both views
    href = {}
    for key, img, dimg, est, b_est, conf, b_conf in (("l", L, dl, oce, b_ce, ocl, b_cl), ("r", Rt, dr, rw, b_rw, ocr, b_cr)):
        b_v, b_h = ctx.alloc(8 * n), ctx.alloc(8 * n)
        ctx.asw_ref_v(w, h, p, dimg.ptr, b_est.ptr, b_conf.ptr, b_v.ptr)
        ov = oracle.asw_ref_v(img, est, conf, radius=R, dscale=dscale, use_fma=True)
        assert_bit_equal(b_v.download((2, h, w), np.float32), ov, f"asw_ref_v {key}")
        ctx.asw_ref_h(w, h, p, dimg.ptr, b_conf.ptr, b_v.ptr, b_h.ptr)
        oh = oracle.asw_ref_h(img, conf, ov, radius=R, use_fma=True)
        assert_bit_equal(b_h.download((2, h, w), np.float32), oh, f"asw_ref_h {key}")
        href[key] = (b_h, oh)
        # the separately rounded arithmetic stays within the north_star tolerance
        oh0 = oracle.asw_ref_h(img, conf, oracle.asw_ref_v(img, est, conf, radius=R, dscale=dscale, use_fma=False), radius=R, use_fma=False)
        assert np.allclose(oh, oh0, rtol=1e-5, atol=0)

    # asw_WTA_REF (asw_wta_ref.cl) on a smooth random volume
    cost = (rng.random((D, h, w), dtype=np.float32) * 40 + 1).astype(np.float32)
    b_cost = ctx.to_device(cost)
    o_l, o_r, f0, f1, f2 = (ctx.alloc(4 * n) for _ in range(5))
    sentinel = np.full((h, w), -7.0, np.float32)
    f3 = ctx.to_device(sentinel)
    ctx.asw_WTA_REF(w, h, p, b_cost.ptr, href["l"][0].ptr, href["r"][0].ptr, o_l.ptr, o_r.ptr, f0.ptr, f1.ptr, f2.ptr, f3.ptr)
    ow = oracle.asw_wta_ref(cost, href["l"][1], href["r"][1], use_fma=True)
    assert_bit_equal(o_l.download((h, w, 4), np.uint8), ow["left"], "WTA_REF output")
    assert_bit_equal(o_r.download((h, w, 4), np.uint8), ow["right"], "WTA_REF output_target")
    assert_bit_equal(f0.download((h, w), np.float32), ow["d_ref"], "WTA_REF disp_ref")
    assert_bit_equal(f1.download((h, w), np.float32), ow["d_tar"], "WTA_REF disp_ref_target")
    assert_bit_equal(f2.download((h, w), np.float32), ow["confidence"], "WTA_REF confidence")
    assert_bit_equal(f3.download((h, w), np.float32), sentinel, "WTA_REF leaves confidence_target untouched")

    # Median (median.cl)
    b_med = ctx.alloc(4 * n)
    ctx.asw_Median(w, h, b_ce.ptr, b_med.ptr)
    assert_bit_equal(b_med.download((h, w, 4), np.uint8), oracle.median(oce), "Median")
    noise = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    b_noise = ctx.to_device(noise)
    ctx.asw_Median(w, h, b_noise.ptr, b_med.ptr)
    assert_bit_equal(b_med.download((h, w, 4), np.uint8), oracle.median(noise), "Median (noise)")


@pytest.mark.parametrize("ds,D,k", [("tsukuba", 61, 6), ("sukub", 16, 2), ("teddy", 61, 6)])
def test_whole_method_parity(ctx, oracle, ds, D, k):
    L, R = load_pair(ds)
    p = P(ndisp=D)
    got = ctx.stereo(L, R, p, refine_iters=k)
    want = oracle.asw_full(L, R, OP(p), use_fma=True, refine_iters=k)
    for key in ("pre_red", "post_red", "disparity"):
        assert_bit_equal(got[key], want[key], f"asw_stereo {key}")
    assert got["timing"]["kernel_launches"] > 0
    tt = got["tail_timing"]
    assert tt["total_ms"] > got["timing"]["total_ms"] > 0 and tt["median_ms"] > 0 and tt["consistency_ms"] > 0
    assert (tt["refinement_total_ms"] > 0) == (k > 0)


def test_whole_method_no_refinement_and_crop(ctx, oracle):
    L, R = crop_pair("cones", 13, 40, 200, 90)
    p = P(ndisp=128, iterations=2)                      # the TMA kernel family (Dp % 128 == 0)
    for k in (0, 1):
        got = ctx.stereo(L, R, p, refine_iters=k)
        want = oracle.asw_full(L, R, OP(p), use_fma=True, refine_iters=k)
        for key in ("pre_red", "post_red", "disparity"):
            assert_bit_equal(got[key], want[key], f"asw_stereo k={k} {key}")


@pytest.mark.parametrize("ds", ["tsukuba", "laundry"])
def test_whole_method_vs_reference_disparity_png(ctx, ds):
    """The shipped path against the PNGs the reference committed (no oracle in between)."""
    L, R = load_pair(ds)
    got = ctx.stereo(L, R)
    g = load_rgba(os.path.join(GOLDEN, ds, "asw_disparity.png"))
    pct = 100.0 * float((g != got["disparity"]).any(-1).mean())
    assert pct <= (0.0 if ds == "tsukuba" else 0.01), f"{ds}: asw_disparity.png differs on {pct:.4f} % of the pixels"
    gp = load_rgba(os.path.join(GOLDEN, ds, "asw_consistency_post-reff.png"))
    assert 100.0 * float((gp != got["post_red"]).any(-1).mean()) <= (0.0 if ds == "tsukuba" else 0.01)


def test_host_binary_whole_method(tmp_path):
    """The pics.txt-driven host program (src/host, C++ on the C ABI) reproduces the reference's PNGs."""
    import shutil
    import subprocess
    from conftest import PAIRS
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "src", "host", "stereo_matching")
    assert os.path.exists(exe), "run `python -m stereo_matchin_b200.build` first"
    os.makedirs(tmp_path / "tsukuba")
    for f in PAIRS["tsukuba"]:
        shutil.copy(os.path.join(GOLDEN, "tsukuba", f), tmp_path / "tsukuba" / f)
    (tmp_path / "pics.txt").write_text("tsukuba/im1.png\ntsukuba/im5.png\n")
    r = subprocess.run([exe, "--pics", str(tmp_path / "pics.txt"), "--root", str(tmp_path), "--runs", "2", "--method", "whole",
                        "--out-suffix", "", "--log", str(tmp_path / "log.tsv")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    for name in ("asw_disparity.png", "asw_consistency_pre-reff.png", "asw_consistency_post-reff.png"):
        assert np.array_equal(load_rgba(str(tmp_path / "tsukuba" / name)), load_rgba(os.path.join(GOLDEN, "tsukuba", name))), name
    log = (tmp_path / "log.tsv").read_text()
    assert "total WTA method" in log and "Run 2" in log
    run1 = [ln for ln in log.splitlines() if ln.startswith("Run 1")][0].split("\t")
    vals = [float(v) for v in run1[1:] if v.strip()]
    assert len(vals) == 14 + 6 + 9 + 1                     # cross-based (0) + hot path + tail + total, as in main.cpp:181
    assert all(v == 0 for v in vals[:14]) and all(v > 0 for v in vals[14:])
