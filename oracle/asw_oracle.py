"""ctypes binding of the CPU oracle (oracle/asw_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from stereo_matchin_b200/.
See the header of asw_oracle.c for scope and parity status (pinned against the
reference's committed PNGs, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libasw_oracle.so")
_lib = None


class _Params(C.Structure):
    _fields_ = [("radius", C.c_int), ("ndisp", C.c_int), ("gamma_c", C.c_float),
                ("gamma_p", C.c_float), ("trunc", C.c_float), ("iterations", C.c_int)]


@dataclass
class OracleParams:
    """Defaults = the reference's literals (asw_vsupport.cl:19,22,24; asw_aggr.cl:16; main.cpp:177)."""
    radius: int = 16
    ndisp: int = 61
    gamma_c: float = 30.91
    gamma_p: float = 28.21
    trunc: float = float("inf")
    iterations: int = 7

    def c(self) -> _Params:
        return _Params(self.radius, self.ndisp, self.gamma_c, self.gamma_p, self.trunc, self.iterations)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("asw_oracle.c", "asw_tail_oracle.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        f32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_num_threads.argtypes = [C.c_int]
        _lib.oracle_q8.restype = C.c_uint8
        _lib.oracle_q8.argtypes = [C.c_float]
        _lib.oracle_asw_aggr.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_float, f32p]
        _lib.oracle_asw_vsupport.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p]
        _lib.oracle_asw_hsupport.argtypes = _lib.oracle_asw_vsupport.argtypes
        _lib.oracle_asw_vcost_aggregation.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p]
        _lib.oracle_asw_hcost_aggregation.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        _lib.oracle_asw_wta.argtypes = [f32p, C.c_int, C.c_int, C.c_int, u8p, f32p, f32p, u8p, f32p, f32p]
        _lib.oracle_consistency.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_float, f32p, f32p, u8p, u8p]
        _lib.oracle_asw_ref_v.argtypes = [u8p, u8p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, f32p]
        _lib.oracle_asw_ref_h.argtypes = [u8p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        _lib.oracle_asw_wta_ref.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, u8p, f32p, f32p, f32p]
        _lib.oracle_median.argtypes = [u8p, C.c_int, C.c_int, u8p]
        _lib.oracle_asw_full.restype = C.c_int
        _lib.oracle_asw_full.argtypes = [u8p, u8p, C.c_int, C.c_int, C.POINTER(_Params), C.c_int, C.c_int, u8p, u8p, u8p]
        i32p = C.c_void_p
        _lib.oracle_cb_median_grid.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        _lib.oracle_cb_cross.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p]
        _lib.oracle_cb_aggregation.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, f32p]
        _lib.oracle_cb_integral_h.argtypes = [f32p, C.c_int, C.c_int, C.c_int]
        _lib.oracle_cb_integral_v.argtypes = [f32p, C.c_int, C.c_int, C.c_int]
        _lib.oracle_cb_oii_hcross.argtypes = [i32p, i32p, f32p, C.c_int, C.c_int, C.c_int, f32p]
        _lib.oracle_cb_oii_vcross.argtypes = [i32p, i32p, f32p, C.c_int, C.c_int, C.c_int, f32p]
        _lib.oracle_cb_init_disparity.argtypes = [f32p, C.c_int, C.c_int, C.c_int, u8p]
        _lib.oracle_cb_disparity.argtypes = [u8p, i32p, C.c_int, C.c_int, C.c_int, u8p]
        _lib.oracle_cross_full.restype = C.c_int
        _lib.oracle_cross_full.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, u8p, u8p]
        _lib.oracle_asw_hot_path.restype = C.c_int
        _lib.oracle_asw_hot_path.argtypes = [u8p, u8p, C.c_int, C.c_int, C.POINTER(_Params), C.c_int,
                                             f32p, u8p, u8p, f32p, f32p, f32p, f32p]
    return _lib


def _f32(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint8))


def _img(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 4, "RGBA8 image expected (H, W, 4)"
    return a


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def q8(f: float) -> int:
    return int(lib().oracle_q8(float(f)))


def asw_aggr(left: np.ndarray, right: np.ndarray, ndisp: int = 61, trunc: float = float("inf")) -> np.ndarray:
    left, right = _img(left), _img(right)
    H, W, _ = left.shape
    out = np.empty((ndisp, H, W), np.float32)
    lib().oracle_asw_aggr(_u8(left), _u8(right), W, H, ndisp, trunc, _f32(out))
    return out


def asw_support(img: np.ndarray, vertical: bool, radius: int = 16, gamma_c: float = 30.91, gamma_p: float = 28.21) -> np.ndarray:
    img = _img(img)
    H, W, _ = img.shape
    out = np.empty((2 * radius + 1, H, W), np.float32)
    fn = lib().oracle_asw_vsupport if vertical else lib().oracle_asw_hsupport
    fn(_u8(img), W, H, radius, gamma_c, gamma_p, _f32(out))
    return out


def asw_vcost_aggregation(supp_l, supp_r, cost, use_fma: bool = True):
    D, H, W = cost.shape
    R = (supp_l.shape[0] - 1) // 2
    cost = np.ascontiguousarray(cost, np.float32)
    out, den = np.empty_like(cost), np.empty_like(cost)
    lib().oracle_asw_vcost_aggregation(_f32(supp_l), _f32(supp_r), _f32(cost), W, H, D, R, int(use_fma), _f32(den), _f32(out))
    return out, den


def asw_hcost_aggregation(supp_l, supp_r, cost, use_fma: bool = True):
    D, H, W = cost.shape
    R = (supp_l.shape[0] - 1) // 2
    cost = np.ascontiguousarray(cost, np.float32)
    out = np.empty_like(cost)
    lib().oracle_asw_hcost_aggregation(_f32(supp_l), _f32(supp_r), _f32(cost), W, H, D, R, int(use_fma), _f32(out))
    return out


def asw_wta(cost: np.ndarray, right_view: bool = True) -> dict:
    cost = np.ascontiguousarray(cost, np.float32)
    D, H, W = cost.shape
    r = {"left": np.empty((H, W, 4), np.uint8), "d_ref": np.empty((H, W), np.float32),
         "conf_ref": np.empty((H, W), np.float32)}
    if right_view:
        r.update(right=np.empty((H, W, 4), np.uint8), d_tar=np.empty((H, W), np.float32),
                 conf_tar=np.empty((H, W), np.float32))
    lib().oracle_asw_wta(_f32(cost), W, H, D, _u8(r["left"]), _f32(r["d_ref"]), _f32(r.get("d_tar")),
                         _u8(r.get("right")), _f32(r["conf_ref"]), _f32(r.get("conf_tar")))
    return r


def consistency(ref_rgba, tar_rgba, conf_ref=None, conf_tar=None, dscale: float = 60.0):
    ref_rgba, tar_rgba = _img(ref_rgba), _img(tar_rgba)
    H, W, _ = ref_rgba.shape
    out, red = np.empty_like(ref_rgba), np.empty_like(ref_rgba)
    lib().oracle_consistency(_u8(ref_rgba), _u8(tar_rgba), W, H, dscale, _f32(conf_ref), _f32(conf_tar), _u8(out), _u8(red))
    return out, red


def asw_hot_path(left: np.ndarray, right: np.ndarray, params: OracleParams | None = None, use_fma: bool = True,
                 want_cost: bool = False, right_view: bool = False) -> dict:
    """raw cost -> 4 support tables -> r x (V, H) -> WTA  (main.cpp:463-526)."""
    params = params or OracleParams()
    left, right = _img(left), _img(right)
    H, W, _ = left.shape
    D = params.ndisp
    r = {"left": np.empty((H, W, 4), np.uint8), "d_ref": np.empty((H, W), np.float32),
         "conf_ref": np.empty((H, W), np.float32)}
    if want_cost:
        r["cost"] = np.empty((D, H, W), np.float32)
    if right_view:
        r.update(right=np.empty((H, W, 4), np.uint8), d_tar=np.empty((H, W), np.float32),
                 conf_tar=np.empty((H, W), np.float32))
    p = params.c()
    rc = lib().oracle_asw_hot_path(_u8(left), _u8(right), W, H, C.byref(p), int(use_fma), _f32(r.get("cost")),
                                   _u8(r["left"]), _u8(r.get("right")), _f32(r["d_ref"]), _f32(r.get("d_tar")),
                                   _f32(r["conf_ref"]), _f32(r.get("conf_tar")))
    if rc != 0:
        raise RuntimeError(f"oracle_asw_hot_path failed rc={rc}")
    return r


# ---- consumers of the hot path (asw_tail_oracle.c) -------------------------------------------------

def asw_ref_v(img, est_rgba, conf, radius: int = 16, dscale: float = 60.0, use_fma: bool = False) -> np.ndarray:
    img, est_rgba = _img(img), _img(est_rgba)
    H, W, _ = img.shape
    conf = np.ascontiguousarray(conf, np.float32)
    out = np.empty((2, H, W), np.float32)
    lib().oracle_asw_ref_v(_u8(img), _u8(est_rgba), _f32(conf), W, H, radius, dscale, int(use_fma), _f32(out))
    return out


def asw_ref_h(img, conf, vref, radius: int = 16, use_fma: bool = False) -> np.ndarray:
    img = _img(img)
    H, W, _ = img.shape
    conf, vref = np.ascontiguousarray(conf, np.float32), np.ascontiguousarray(vref, np.float32)
    out = np.empty((2, H, W), np.float32)
    lib().oracle_asw_ref_h(_u8(img), _f32(conf), _f32(vref), W, H, radius, int(use_fma), _f32(out))
    return out


def asw_wta_ref(cost, href_l, href_r, use_fma: bool = False) -> dict:
    cost = np.ascontiguousarray(cost, np.float32)
    D, H, W = cost.shape
    r = {"left": np.empty((H, W, 4), np.uint8), "right": np.empty((H, W, 4), np.uint8), "d_ref": np.empty((H, W), np.float32),
         "d_tar": np.empty((H, W), np.float32), "confidence": np.empty((H, W), np.float32)}
    lib().oracle_asw_wta_ref(_f32(cost), _f32(np.ascontiguousarray(href_l, np.float32)), _f32(np.ascontiguousarray(href_r, np.float32)),
                             W, H, D, int(use_fma), _u8(r["left"]), _u8(r["right"]), _f32(r["d_ref"]), _f32(r["d_tar"]), _f32(r["confidence"]))
    return r


def median(img_rgba) -> np.ndarray:
    img_rgba = _img(img_rgba)
    H, W, _ = img_rgba.shape
    out = np.empty_like(img_rgba)
    lib().oracle_median(_u8(img_rgba), W, H, _u8(out))
    return out


def asw_full(left, right, params: OracleParams | None = None, use_fma: bool = False, refine_iters: int = 6) -> dict:
    """The whole ASW method of main.cpp:463-631: hot path, consistency, k refinement rounds, median."""
    params = params or OracleParams()
    left, right = _img(left), _img(right)
    H, W, _ = left.shape
    r = {"disparity": np.empty((H, W, 4), np.uint8), "pre_red": np.empty((H, W, 4), np.uint8), "post_red": np.empty((H, W, 4), np.uint8)}
    p = params.c()
    rc = lib().oracle_asw_full(_u8(left), _u8(right), W, H, C.byref(p), int(use_fma), refine_iters, _u8(r["disparity"]),
                               _u8(r["pre_red"]), _u8(r["post_red"]))
    if rc != 0:
        raise RuntimeError(f"oracle_asw_full failed rc={rc}")
    return r


# ---- the cross-based method (cross_oracle.c) ---------------------------------------------------------

def _i32(a):
    return a.ctypes.data_as(C.c_void_p)


def cb_median_grid(img_rgba, local: int = 3) -> np.ndarray:
    img_rgba = _img(img_rgba)
    H, W, _ = img_rgba.shape
    out = np.empty_like(img_rgba)
    lib().oracle_cb_median_grid(_u8(img_rgba), W, H, local, _u8(out))
    return out


def cb_cross(img_rgba, max_arm: int = 25) -> np.ndarray:
    img_rgba = _img(img_rgba)
    H, W, _ = img_rgba.shape
    out = np.empty((4, H, W), np.int32)
    lib().oracle_cb_cross(_u8(img_rgba), W, H, max_arm, _i32(out))
    return out


def cb_aggregation(left, right, D: int = 61) -> np.ndarray:
    left, right = _img(left), _img(right)
    H, W, _ = left.shape
    cost = np.empty((D, H, W), np.float32)
    lib().oracle_cb_aggregation(_u8(left), _u8(right), W, H, D, _f32(cost))
    return cost


def cb_integral(cost, horizontal: bool) -> np.ndarray:
    out = np.array(cost, np.float32, order="C", copy=True)
    D, H, W = out.shape
    (lib().oracle_cb_integral_h if horizontal else lib().oracle_cb_integral_v)(_f32(out), W, H, D)
    return out


def cb_oii(cross_l, cross_r, integral, horizontal: bool) -> np.ndarray:
    cross_l, cross_r = np.ascontiguousarray(cross_l, np.int32), np.ascontiguousarray(cross_r, np.int32)
    integral = np.ascontiguousarray(integral, np.float32)
    D, H, W = integral.shape
    out = np.empty_like(integral)
    (lib().oracle_cb_oii_hcross if horizontal else lib().oracle_cb_oii_vcross)(_i32(cross_l), _i32(cross_r), _f32(integral), W, H, D, _f32(out))
    return out


def cb_init_disparity(cost) -> np.ndarray:
    cost = np.ascontiguousarray(cost, np.float32)
    D, H, W = cost.shape
    out = np.empty((H, W, 4), np.uint8)
    lib().oracle_cb_init_disparity(_f32(cost), W, H, D, _u8(out))
    return out


def cb_disparity(init_rgba, cross, D: int = 61) -> np.ndarray:
    init_rgba, cross = _img(init_rgba), np.ascontiguousarray(cross, np.int32)
    H, W, _ = init_rgba.shape
    out = np.empty_like(init_rgba)
    lib().oracle_cb_disparity(_u8(init_rgba), _i32(cross), W, H, D, _u8(out))
    return out


def cross_full(left, right, D: int = 61, max_arm: int = 25, median_local: int = 3) -> dict:
    """The whole cross-based method of main.cpp:258-367."""
    left, right = _img(left), _img(right)
    H, W, _ = left.shape
    r = {k: np.empty((H, W, 4), np.uint8) for k in ("initial", "final", "median_l")}
    rc = lib().oracle_cross_full(_u8(left), _u8(right), W, H, D, max_arm, median_local, _u8(r["initial"]), _u8(r["final"]), _u8(r["median_l"]))
    if rc != 0:
        raise RuntimeError(f"oracle_cross_full failed rc={rc}")
    return r
