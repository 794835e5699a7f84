"""Checks the C oracle operator by operator against a slow, literal numpy/python reading of
the reference kernels (one python loop nest per __kernel, float32 arithmetic, no FMA), on
small crops, including the edge cases the kernels have: clamped taps at every border,
x < d (clamped right column), 1-pixel-wide / 1-pixel-high images, D = 1."""
import numpy as np
import pytest

from conftest import crop_pair

f32 = np.float32


def px(img):  # read_imagef * 255 (asw_aggr.cl:12)
    return (img[..., :3].astype(f32) / f32(255.0)) * f32(255.0)


def ref_aggr(L, R, D):  # asw_aggr.cl:3-23
    H, W, _ = L.shape
    l, r = px(L), px(R)
    out = np.zeros((D, H, W), f32)
    for d in range(D):
        for x in range(W):
            xr = max(x - d, 0)
            a = np.abs(l[:, x] - r[:, xr])
            out[d, :, x] = (a[:, 0] + a[:, 1]) + a[:, 2]
    return out


def ref_support(img, vertical, R=16, gc=30.91, gp=28.21):  # asw_vsupport.cl / asw_hsupport.cl
    H, W, _ = img.shape
    p = px(img)
    out = np.zeros((2 * R + 1, H, W), f32)
    for i in range(2 * R + 1):
        for y in range(H):
            for x in range(W):
                qy, qx = (min(max(y + i - R, 0), H - 1), x) if vertical else (y, min(max(x + i - R, 0), W - 1))
                a = np.abs(p[y, x] - p[qy, qx])
                sad = f32(f32(a[0] + a[1]) + a[2])
                c = f32(-sad) / f32(gc)
                g = f32(abs(y - qy) + abs(x - qx)) / f32(gp)
                out[i, y, x] = f32(np.exp(np.float64(f32(c - g))))
    return out


def ref_agg(sl, sr, cin, vertical, R=16):  # asw_vcost_aggregation.cl:23-43 / asw_hcost_aggregation.cl:24-43
    D, H, W = cin.shape
    out = np.zeros_like(cin)
    den_out = np.zeros_like(cin)
    for d in range(D):
        for y in range(H):
            for x in range(W):
                xr = max(x - d, 0)
                num, den = f32(0.00001), f32(0.00001)
                for i in range(2 * R + 1):
                    ww = f32(sl[i, y, x] * sr[i, y, xr])
                    c = cin[d, min(max(y + i - R, 0), H - 1), x] if vertical else cin[d, y, min(max(x + i - R, 0), W - 1)]
                    num = f32(num + f32(ww * c))
                    den = f32(den + ww)
                out[d, y, x] = num / den
                den_out[d, y, x] = den
    return out, den_out


def ref_wta(cost):  # asw_wta.cl:25-82
    D, H, W = cost.shape
    res = {k: np.zeros((H, W), f32) for k in ("d_ref", "d_tar", "conf_ref", "conf_tar")}
    for y in range(H):
        for x in range(W):
            cur = last = f32(100000)
            md = 0
            for i in range(D):
                t = cost[i, y, x]
                if t < last: last = t
                if t < cur: md = i
                if t < cur: last = cur
                if t < cur: cur = t
            mdr, cur_t, last_t = md, f32(100000), f32(100000)
            for i in range(md):
                xq = max(0, x - i)
                b = xq - x + md   # bresenham((0,x-d),(d,x),xq) = 1*(xq-x)+d, asw_wta.cl:3-9,57
                t = cost[b, y, xq]
                if t < last_t: last_t = t
                if t < cur_t: mdr = b
                if t < cur_t: last_t = cur_t
                if t < cur_t: cur_t = t
            res["d_ref"][y, x], res["d_tar"][y, x] = md, mdr
            res["conf_ref"][y, x] = (last - cur) / last
            res["conf_tar"][y, x] = (last_t - cur_t) / last_t
    return res


CROPS = [("teddy", 0, 0, 24, 9, 12), ("cones", 200, 100, 19, 21, 7), ("tsukuba", 370, 270, 14, 18, 20),
         ("art", 5, 5, 1, 40, 3), ("laundry", 0, 300, 40, 1, 5), ("teddy", 100, 100, 9, 9, 1)]


@pytest.mark.parametrize("ds,x0,y0,w,h,D", CROPS)
def test_operators_match_literal_restatement(oracle, ds, x0, y0, w, h, D):
    L, R = crop_pair(ds, x0, y0, w, h)
    raw = oracle.asw_aggr(L, R, D)
    assert np.array_equal(raw, ref_aggr(L, R, D))
    tabs = {}
    for name, img, vert in (("vl", L, True), ("hl", L, False), ("vr", R, True), ("hr", R, False)):
        tabs[name] = oracle.asw_support(img, vert)
        assert np.array_equal(tabs[name], ref_support(img, vert)), name
        assert np.all(tabs[name][16] == 1.0)   # centre tap is exactly 1
    v, den = oracle.asw_vcost_aggregation(tabs["vl"], tabs["vr"], raw, use_fma=False)
    rv, rden = ref_agg(tabs["vl"], tabs["vr"], raw, True)
    assert np.array_equal(v, rv) and np.array_equal(den, rden)
    h_ = oracle.asw_hcost_aggregation(tabs["hl"], tabs["hr"], v, use_fma=False)
    rh, _ = ref_agg(tabs["hl"], tabs["hr"], v, False)
    assert np.array_equal(h_, rh)
    # the FMA variant differs from the separately rounded one by float noise only
    vf, _ = oracle.asw_vcost_aggregation(tabs["vl"], tabs["vr"], raw, use_fma=True)
    assert np.allclose(vf, v, rtol=1e-5, atol=0)   # north_star: aggregated costs within 1e-5 relative
    wta = oracle.asw_wta(h_)
    rw = ref_wta(h_)
    for k in rw:
        assert np.array_equal(wta[k], rw[k]), k
    # the one-call hot path equals the operator sequence
    hp = oracle.asw_hot_path(L, R, oracle.OracleParams(ndisp=D, iterations=1), use_fma=False, want_cost=True, right_view=True)
    assert np.array_equal(hp["cost"], h_) and np.array_equal(hp["d_tar"], wta["d_tar"])


def test_trunc_and_iterations_zero(oracle):
    L, R = crop_pair("teddy", 50, 50, 16, 8)
    raw = oracle.asw_aggr(L, R, 8)
    cut = oracle.asw_aggr(L, R, 8, trunc=20.0)
    assert np.array_equal(cut, np.minimum(raw, np.float32(20.0)))
    hp = oracle.asw_hot_path(L, R, oracle.OracleParams(ndisp=8, iterations=0), want_cost=True)
    assert np.array_equal(hp["cost"], raw)
    assert np.array_equal(hp["d_ref"], raw.argmin(0).astype(np.float32))


def test_wta_two_min_semantics(oracle):
    """Exact ties: lowest d wins and conf = 0; the second minimum is the second-smallest VALUE."""
    c = np.array([5, 3, 7, 3, 9], np.float32).reshape(5, 1, 1)
    w = oracle.asw_wta(c, right_view=False)
    assert w["d_ref"][0, 0] == 1 and w["conf_ref"][0, 0] == 0.0
    c = np.array([5, 3, 7, 4, 9], np.float32).reshape(5, 1, 1)
    w = oracle.asw_wta(c, right_view=False)
    assert w["d_ref"][0, 0] == 1 and w["conf_ref"][0, 0] == np.float32(np.float32(1) / np.float32(4))
    c = np.array([2], np.float32).reshape(1, 1, 1)   # D = 1: second minimum stays at the sentinel 100000
    w = oracle.asw_wta(c, right_view=False)
    assert w["conf_ref"][0, 0] == np.float32((np.float32(100000) - np.float32(2)) / np.float32(100000))
