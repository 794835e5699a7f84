"""ctypes binding of the C ABI (include/asw_b200.h) -- the host-side mirror of the reference's
operator surface for the ASW hot path (kernels/asw_*.cl + the enqueue sequence of
stereo_matching/main.cpp:463-526).

The product is libasw_b200.so (hand-written CUDA for sm_100a).  This module only marshals
arguments: numpy arrays for host buffers, integer device addresses (e.g. torch
``tensor.data_ptr()``) for device buffers.  There is no CPU fallback: if the library is
missing or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ASW_B200_LIB") or os.path.join(_PKG, "libasw_b200.so")   # override: A/B of experimental builds

ASW_OK, ASW_ERR_INVALID, ASW_ERR_CUDA, ASW_ERR_NOMEM, ASW_ERR_UNSUPPORTED = range(5)

# every symbol include/asw_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "asw_version", "asw_strerror", "asw_create", "asw_destroy", "asw_last_error", "asw_stream", "asw_sync",
    "asw_device_info", "asw_params_default", "asw_disparity", "asw_disparity_async", "asw_disparity_device", "asw_disparity_band_device", "asw_disparity_band_exchange_device", "asw_disparity_band_exchange_async_device", "asw_side_stream",
    "asw_multi_create", "asw_multi_destroy", "asw_multi_count", "asw_multi_last_error", "asw_multi_disparity",
    "asw_disparity_shard_device", "asw_merge_shards",
    "asw_set_keep_volume", "asw_final_volume", "asw_Aggr", "asw_vSupport", "asw_hSupport", "asw_vCostAggregation",
    "asw_hCostAggregation", "asw_WTA", "asw_Constistency", "asw_ref_v", "asw_ref_h", "asw_WTA_REF", "asw_Median", "asw_stereo",
    "asw_cross_params_default", "asw_Median_grid", "asw_Cross", "asw_Aggregation", "asw_Integral_h", "asw_Integral_v", "asw_Oii_hcross",
    "asw_Oii_vcross", "asw_Init_disparity", "asw_Disparity", "asw_cross_stereo",
    "asw_dev_alloc", "asw_dev_free", "asw_memcpy_h2d", "asw_memcpy_d2h",
    "asw_host_alloc", "asw_host_free", "asw_set_kernel_family",
]


class AswError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"asw_b200 status {status}: {msg}")
        self.status = status


class CParams(C.Structure):
    _fields_ = [("radius", C.c_int), ("ndisp", C.c_int), ("gamma_c", C.c_float), ("gamma_p", C.c_float),
                ("trunc", C.c_float), ("iterations", C.c_int)]


class CTiming(C.Structure):
    _fields_ = [("raw_ms", C.c_float), ("supp_ms", C.c_float), ("vagg_mean_ms", C.c_float), ("hagg_mean_ms", C.c_float),
                ("agg_total_ms", C.c_float), ("wta_ms", C.c_float), ("total_ms", C.c_float), ("h2d_ms", C.c_float),
                ("d2h_ms", C.c_float), ("kernel_launches", C.c_int), ("vfix_mean_ms", C.c_float)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class CMultiTiming(C.Structure):
    _fields_ = [("devices", C.c_int), ("upload_ms", C.c_float), ("compute_ms", C.c_float), ("download_ms", C.c_float),
                ("total_ms", C.c_float), ("slowest_band_device_ms", C.c_float)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


# asw_halo_fn: int (*)(void* user, int iteration, void* top_send, void* bottom_send, void* top_recv, void* bottom_recv, size_t bytes)
HALO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)


# asw_halo_begin_fn / asw_halo_end_fn (asynchronous exchange hidden under the interior rows)
HALO_BEGIN_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
HALO_END_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)


class CTailTiming(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("right_wta_ms", "consistency_ms", "vref_mean_l_ms", "vref_mean_r_ms", "href_mean_l_ms",
                                          "href_mean_r_ms", "wta_ref_mean_ms", "consistency_mean_ms", "refinement_total_ms",
                                          "median_ms", "total_ms")]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class CCrossParams(C.Structure):
    _fields_ = [("ndisp", C.c_int), ("max_arm", C.c_int), ("median_local", C.c_int)]


class CCrossTiming(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("median_l_ms", "median_r_ms", "median_ms", "cross_l_ms", "cross_r_ms", "cross_ms", "aggregation_ms",
                                          "integral_h_ms", "oii_h_ms", "integral_v_ms", "oii_v_ms", "init_disparity_ms",
                                          "final_disparity_ms", "total_ms")]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


@dataclass
class CrossParams:
    """Defaults = the reference's literals (aggregation.cl:14, cross.cl:33-81, main.cpp:191-197)."""
    ndisp: int = 61
    max_arm: int = 25
    median_local: int = 3

    def c(self) -> CCrossParams:
        return CCrossParams(self.ndisp, self.max_arm, self.median_local)


@dataclass
class AswParams:
    """Defaults reproduce the reference's literals (asw_vsupport.cl:19,22,24; asw_aggr.cl:16; main.cpp:177)."""
    radius: int = 16
    ndisp: int = 61
    gamma_c: float = 30.91
    gamma_p: float = 28.21
    trunc: float = math.inf
    iterations: int = 7

    def c(self) -> CParams:
        return CParams(self.radius, self.ndisp, self.gamma_c, self.gamma_p, self.trunc, self.iterations)


_lib = None


def load_library() -> C.CDLL:
    """Loads libasw_b200.so (never builds, never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -m stereo_matchin_b200.build` (needs nvcc)")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, f32p, ip = C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)
    pp, tp = C.POINTER(CParams), C.POINTER(CTiming)
    lib.asw_version.restype = C.c_char_p
    lib.asw_strerror.restype = C.c_char_p
    lib.asw_strerror.argtypes = [C.c_int]
    lib.asw_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.asw_destroy.argtypes = [vp]
    lib.asw_last_error.restype = C.c_char_p
    lib.asw_last_error.argtypes = [vp]
    lib.asw_stream.restype = vp
    lib.asw_stream.argtypes = [vp]
    lib.asw_sync.argtypes = [vp]
    lib.asw_device_info.argtypes = [vp, ip, ip, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    lib.asw_params_default.argtypes = [pp]
    lib.asw_params_default.restype = None
    lib.asw_disparity.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, pp, u8p, u8p, f32p, tp]
    lib.asw_disparity_async.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, pp, u8p, u8p, f32p]
    lib.asw_disparity_device.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, pp, u8p, u8p, f32p, tp]
    lib.asw_disparity_band_device.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, pp, u8p, u8p, f32p, tp]
    lib.asw_disparity_band_exchange_device.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, pp, u8p, u8p, f32p, HALO_FN, vp, tp]
    lib.asw_disparity_band_exchange_async_device.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, pp, u8p, u8p, f32p, HALO_BEGIN_FN,
                                                             HALO_END_FN, vp, tp]
    lib.asw_side_stream.restype = vp
    lib.asw_side_stream.argtypes = [vp]
    lib.asw_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    lib.asw_multi_destroy.argtypes = [vp]
    lib.asw_multi_count.argtypes = [vp]
    lib.asw_multi_last_error.restype = C.c_char_p
    lib.asw_multi_last_error.argtypes = [vp]
    lib.asw_multi_disparity.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, pp, u8p, u8p, f32p, C.POINTER(CMultiTiming)]
    lib.asw_disparity_shard_device.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, pp, f32p, f32p, vp, tp]
    lib.asw_merge_shards.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, vp, u8p, u8p, f32p]
    lib.asw_set_keep_volume.argtypes = [vp, C.c_int]
    lib.asw_final_volume.restype = vp
    lib.asw_final_volume.argtypes = [vp]
    lib.asw_Aggr.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, pp, f32p]
    lib.asw_vSupport.argtypes = [vp, u8p, C.c_int, C.c_int, pp, f32p]
    lib.asw_hSupport.argtypes = [vp, u8p, C.c_int, C.c_int, pp, f32p]
    lib.asw_vCostAggregation.argtypes = [vp, C.c_int, C.c_int, pp, f32p, f32p, f32p, f32p, f32p]
    lib.asw_hCostAggregation.argtypes = [vp, C.c_int, C.c_int, pp, f32p, f32p, f32p, f32p, f32p]
    lib.asw_WTA.argtypes = [vp, C.c_int, C.c_int, pp, f32p, u8p, f32p, f32p, u8p, f32p, f32p]
    lib.asw_Constistency.argtypes = [vp, C.c_int, C.c_int, pp, u8p, u8p, f32p, f32p, u8p, u8p]
    lib.asw_ref_v.argtypes = [vp, C.c_int, C.c_int, pp, u8p, u8p, f32p, f32p]
    lib.asw_ref_h.argtypes = [vp, C.c_int, C.c_int, pp, u8p, f32p, f32p, f32p]
    lib.asw_WTA_REF.argtypes = [vp, C.c_int, C.c_int, pp, f32p, f32p, f32p, u8p, u8p, f32p, f32p, f32p, f32p]
    lib.asw_Median.argtypes = [vp, C.c_int, C.c_int, u8p, u8p]
    lib.asw_stereo.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, pp, C.c_int, u8p, u8p, u8p, tp, C.POINTER(CTailTiming)]
    cp = C.POINTER(CCrossParams)
    lib.asw_cross_params_default.argtypes = [cp]
    lib.asw_cross_params_default.restype = None
    lib.asw_Median_grid.argtypes = [vp, C.c_int, C.c_int, C.c_int, u8p, u8p]
    lib.asw_Cross.argtypes = [vp, C.c_int, C.c_int, cp, u8p, vp]
    lib.asw_Aggregation.argtypes = [vp, C.c_int, C.c_int, cp, u8p, u8p, f32p]
    lib.asw_Integral_h.argtypes = [vp, C.c_int, C.c_int, cp, f32p]
    lib.asw_Integral_v.argtypes = [vp, C.c_int, C.c_int, cp, f32p]
    lib.asw_Oii_hcross.argtypes = [vp, C.c_int, C.c_int, cp, vp, vp, f32p, f32p]
    lib.asw_Oii_vcross.argtypes = [vp, C.c_int, C.c_int, cp, vp, vp, f32p, f32p]
    lib.asw_Init_disparity.argtypes = [vp, C.c_int, C.c_int, cp, f32p, u8p]
    lib.asw_Disparity.argtypes = [vp, C.c_int, C.c_int, cp, u8p, vp, u8p]
    lib.asw_cross_stereo.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, cp, u8p, u8p, u8p, C.POINTER(CCrossTiming)]
    lib.asw_dev_alloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    lib.asw_dev_free.argtypes = [vp, vp]
    lib.asw_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    lib.asw_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    lib.asw_host_alloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    lib.asw_host_free.argtypes = [vp, vp]
    lib.asw_set_kernel_family.argtypes = [vp, C.c_int]
    _lib = lib
    return lib


def default_params() -> AswParams:
    p = CParams()
    load_library().asw_params_default(C.byref(p))
    return AswParams(p.radius, p.ndisp, p.gamma_c, p.gamma_p, p.trunc, p.iterations)


def _host_ptr(a):
    return None if a is None else a.ctypes.data


def _rgba(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("RGBA8 image of shape (H, W, 4) expected")
    return a


class DeviceBuffer:
    """A device allocation owned through the ABI (asw_dev_alloc / asw_dev_free)."""

    def __init__(self, ctx: "AswContext", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        ctx._check(ctx.lib.asw_dev_alloc(ctx.h, C.byref(p), self.nbytes))
        self.ptr = p.value

    def upload(self, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.asw_memcpy_h2d(self.ctx.h, self.ptr, a.ctypes.data, a.nbytes))
        return self

    def download(self, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.asw_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.asw_dev_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class AswContext:
    """One GPU + one stream (the reference's context + in-order queue, main.cpp:170-171,212)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.asw_create(C.byref(h), int(device))
        if st != ASW_OK:
            raise AswError(st, f"asw_create(device={device}) failed: {self.lib.asw_strerror(st).decode()} "
                               "(a CUDA device is required; there is no CPU fallback)")
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.asw_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != ASW_OK:
            raise AswError(st, f"{self.lib.asw_strerror(st).decode()}: {self.lib.asw_last_error(self.h).decode()}")

    # -- misc ------------------------------------------------------------------------------------
    def stream(self) -> int:
        return int(self.lib.asw_stream(self.h) or 0)

    def sync(self):
        self._check(self.lib.asw_sync(self.h))

    def device_info(self) -> dict:
        sm, khz, mem = C.c_int(), C.c_int(), C.c_size_t()
        name = C.create_string_buffer(256)
        self._check(self.lib.asw_device_info(self.h, C.byref(sm), C.byref(khz), C.byref(mem), name, 256))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "total_mem": mem.value, "name": name.value.decode()}

    def set_kernel_family(self, family: int):
        self._check(self.lib.asw_set_kernel_family(self.h, family))

    def set_keep_volume(self, keep: bool):
        self._check(self.lib.asw_set_keep_volume(self.h, int(keep)))

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def to_device(self, a: np.ndarray) -> DeviceBuffer:
        a = np.ascontiguousarray(a)
        return DeviceBuffer(self, a.nbytes).upload(a)

    # -- fused hot path ----------------------------------------------------------------------------
    def disparity(self, left: np.ndarray, right: np.ndarray, params: AswParams | None = None, want_conf: bool = True,
                  want_timing: bool = False) -> dict:
        """Host buffers in, host buffers out (asw_disparity): the call a user of the reference makes."""
        params = params or AswParams()
        left, right = _rgba(left), _rgba(right)
        if left.shape != right.shape:
            raise ValueError("left and right images must have the same shape")
        H, W, _ = left.shape
        out = {"disp_rgba": np.empty((H, W, 4), np.uint8),
               "disp_d": np.empty((H, W), np.uint8) if params.ndisp <= 256 else None,
               "conf": np.empty((H, W), np.float32) if want_conf else None}
        tm = CTiming() if want_timing else None
        p = params.c()
        self._check(self.lib.asw_disparity(self.h, left.ctypes.data, right.ctypes.data, W, H, C.byref(p),
                                           _host_ptr(out["disp_rgba"]), _host_ptr(out["disp_d"]), _host_ptr(out["conf"]),
                                           C.byref(tm) if tm is not None else None))
        if tm is not None:
            out["timing"] = tm.as_dict()
        return out

    def disparity_raw(self, left_ptr: int, right_ptr: int, W: int, H: int, params: AswParams, rgba_ptr: int | None,
                      d_ptr: int | None, conf_ptr: int | None, timing: bool = False, host: bool = False,
                      band: tuple[int, int] | None = None) -> dict | None:
        """Pointer-level call: host pointers (host=True, asw_disparity) or device pointers."""
        tm = CTiming() if timing else None
        p = params.c()
        tref = C.byref(tm) if tm is not None else None
        if host:
            st = self.lib.asw_disparity(self.h, left_ptr, right_ptr, W, H, C.byref(p), rgba_ptr, d_ptr, conf_ptr, tref)
        elif band is not None:
            st = self.lib.asw_disparity_band_device(self.h, left_ptr, right_ptr, W, H, band[0], band[1], C.byref(p), rgba_ptr,
                                                    d_ptr, conf_ptr, tref)
        else:
            st = self.lib.asw_disparity_device(self.h, left_ptr, right_ptr, W, H, C.byref(p), rgba_ptr, d_ptr, conf_ptr, tref)
        self._check(st)
        return tm.as_dict() if tm is not None else None

    def disparity_band_exchange(self, left_ptr: int, right_ptr: int, W: int, H: int, params: AswParams, band: tuple[int, int],
                                rgba_ptr: int | None, d_ptr: int | None, conf_ptr: int | None, exchange, timing: bool = False):
        """asw_disparity_band_exchange_device: rows band = (y0, y1) with a per-iteration halo exchange.
        `exchange(iteration, top_send, bottom_send, top_recv, bottom_recv, nbytes)` receives device addresses (None towards
        a frame border) and must have moved the rows when it returns (see include/asw_b200.h for the ordering contract)."""
        err = []

        def _cb(_user, it, ts, bs, tr, br, nbytes):
            try:
                exchange(it, ts, bs, tr, br, nbytes)
                return 0
            except Exception as e:      # an exception must not cross the C frames
                err.append(e)
                return 1

        cb = HALO_FN(_cb)
        tm = CTiming() if timing else None
        p = params.c()
        st = self.lib.asw_disparity_band_exchange_device(self.h, left_ptr, right_ptr, W, H, band[0], band[1], C.byref(p), rgba_ptr, d_ptr,
                                                         conf_ptr, cb, None, C.byref(tm) if tm is not None else None)
        if err:
            raise err[0]
        self._check(st)
        return tm.as_dict() if tm is not None else None

    def disparity_band_exchange_async(self, left_ptr: int, right_ptr: int, W: int, H: int, params: AswParams, band: tuple[int, int],
                                      rgba_ptr: int | None, d_ptr: int | None, conf_ptr: int | None, begin, end, timing: bool = False):
        """asw_disparity_band_exchange_async_device: `begin(iteration, top_send, bottom_send, top_recv, bottom_recv, nbytes,
        boundary_stream)` starts the transfer (ordered after `boundary_stream`) and returns; `end(iteration, main_stream)` makes
        `main_stream` wait for it.  Streams are raw cudaStream_t addresses.  Neither may block on another band's GPU work."""
        err = []

        def _b(_u, it, ts, bs, tr, br, nbytes, st):
            try:
                begin(it, ts, bs, tr, br, nbytes, st)
                return 0
            except Exception as e:
                err.append(e)
                return 1

        def _e(_u, it, st):
            try:
                end(it, st)
                return 0
            except Exception as e:
                err.append(e)
                return 1

        cb, ce = HALO_BEGIN_FN(_b), HALO_END_FN(_e)
        tm = CTiming() if timing else None
        p = params.c()
        st = self.lib.asw_disparity_band_exchange_async_device(self.h, left_ptr, right_ptr, W, H, band[0], band[1], C.byref(p), rgba_ptr,
                                                               d_ptr, conf_ptr, cb, ce, None, C.byref(tm) if tm is not None else None)
        if err:
            raise err[0]
        self._check(st)
        return tm.as_dict() if tm is not None else None

    def disparity_async(self, left_ptr: int, right_ptr: int, W: int, H: int, params: AswParams, rgba_ptr: int | None,
                        d_ptr: int | None, conf_ptr: int | None):
        """asw_disparity_async: pinned HOST pointers; upload + hot path + download are enqueued, `sync()` completes them."""
        p = params.c()
        self._check(self.lib.asw_disparity_async(self.h, left_ptr, right_ptr, W, H, C.byref(p), rgba_ptr, d_ptr, conf_ptr))

    def stereo(self, left: np.ndarray, right: np.ndarray, params: AswParams | None = None, refine_iters: int = 6) -> dict:
        """The whole ASW method (asw_stereo): final disparity image + the two consistency images."""
        params = params or AswParams()
        left, right = _rgba(left), _rgba(right)
        H, W, _ = left.shape
        out = {k: np.empty((H, W, 4), np.uint8) for k in ("disparity", "pre_red", "post_red")}
        tm, tt = CTiming(), CTailTiming()
        p = params.c()
        self._check(self.lib.asw_stereo(self.h, left.ctypes.data, right.ctypes.data, W, H, C.byref(p), refine_iters,
                                        out["disparity"].ctypes.data, out["pre_red"].ctypes.data, out["post_red"].ctypes.data, C.byref(tm),
                                        C.byref(tt)))
        out["timing"] = tm.as_dict()
        out["tail_timing"] = tt.as_dict()
        return out

    def cross_stereo(self, left: np.ndarray, right: np.ndarray, params: "CrossParams | None" = None) -> dict:
        """The whole cross-based method (asw_cross_stereo): initial and final disparity images + the median of the left image."""
        params = params or CrossParams()
        left, right = _rgba(left), _rgba(right)
        H, W, _ = left.shape
        out = {k: np.empty((H, W, 4), np.uint8) for k in ("initial", "final", "median_l")}
        tm, p = CCrossTiming(), params.c()
        self._check(self.lib.asw_cross_stereo(self.h, left.ctypes.data, right.ctypes.data, W, H, C.byref(p), out["initial"].ctypes.data,
                                              out["final"].ctypes.data, out["median_l"].ctypes.data, C.byref(tm)))
        out["timing"] = tm.as_dict()
        return out

    def cb_op(self, name: str, W: int, H: int, params: "CrossParams", *ptrs):
        """Per-operator entry points of the cross-based method: asw_Cross, asw_Aggregation, asw_Integral_h, ..."""
        p = params.c()
        self._check(getattr(self.lib, name)(self.h, W, H, C.byref(p), *ptrs))

    def asw_Median_grid(self, W, H, local, input_, output):
        self._check(self.lib.asw_Median_grid(self.h, W, H, local, input_, output))

    def asw_Constistency(self, W, H, params, ref, tar, confidence_ref, confidence_tar, output, output_red):
        p = params.c()
        self._check(self.lib.asw_Constistency(self.h, W, H, C.byref(p), ref, tar, confidence_ref, confidence_tar, output, output_red))

    def asw_ref_v(self, W, H, params, input_, input_est, confidence, output_REF):
        p = params.c()
        self._check(self.lib.asw_ref_v(self.h, W, H, C.byref(p), input_, input_est, confidence, output_REF))

    def asw_ref_h(self, W, H, params, input_, confidence, input_REF, output_REF):
        p = params.c()
        self._check(self.lib.asw_ref_h(self.h, W, H, C.byref(p), input_, confidence, input_REF, output_REF))

    def asw_WTA_REF(self, W, H, params, agg_d, ref, ref_target, output, output_target, disp_ref, disp_ref_target, confidence,
                    confidence_target):
        p = params.c()
        self._check(self.lib.asw_WTA_REF(self.h, W, H, C.byref(p), agg_d, ref, ref_target, output, output_target, disp_ref,
                                         disp_ref_target, confidence, confidence_target))

    def asw_Median(self, W, H, input_, output):
        self._check(self.lib.asw_Median(self.h, W, H, input_, output))

    def disparity_shard_raw(self, left_ptr, right_ptr, W, H, params, band, dshard, min1_ptr, min2_ptr, arg_ptr, timing=False):
        """asw_disparity_shard_device: rows band = (y0, y1), disparities dshard = (d0, d1); partial WTA triples out."""
        tm = CTiming() if timing else None
        p = params.c()
        self._check(self.lib.asw_disparity_shard_device(self.h, left_ptr, right_ptr, W, H, band[0], band[1], dshard[0], dshard[1], C.byref(p),
                                                        min1_ptr, min2_ptr, arg_ptr, C.byref(tm) if tm is not None else None))
        return tm.as_dict() if tm is not None else None

    def merge_shards(self, W, rows, ndisp, nshards, min1_ptr, min2_ptr, arg_ptr, rgba_ptr, d_ptr, conf_ptr):
        self._check(self.lib.asw_merge_shards(self.h, W, rows, ndisp, nshards, min1_ptr, min2_ptr, arg_ptr, rgba_ptr, d_ptr, conf_ptr))

    def final_volume_ptr(self) -> int:
        return int(self.lib.asw_final_volume(self.h) or 0)

    # -- per-operator entry points (device pointers, reference layouts) ----------------------------
    def asw_Aggr(self, input_l: int, input_r: int, W: int, H: int, params: AswParams, output_cost: int):
        p = params.c()
        self._check(self.lib.asw_Aggr(self.h, input_l, input_r, W, H, C.byref(p), output_cost))

    def asw_vSupport(self, input_: int, W: int, H: int, params: AswParams, output: int):
        p = params.c()
        self._check(self.lib.asw_vSupport(self.h, input_, W, H, C.byref(p), output))

    def asw_hSupport(self, input_: int, W: int, H: int, params: AswParams, output: int):
        p = params.c()
        self._check(self.lib.asw_hSupport(self.h, input_, W, H, C.byref(p), output))

    def asw_vCostAggregation(self, W: int, H: int, params: AswParams, supp_left: int, supp_right: int, input_cost: int,
                             output_denom: int | None, output_cost: int):
        p = params.c()
        self._check(self.lib.asw_vCostAggregation(self.h, W, H, C.byref(p), supp_left, supp_right, input_cost, output_denom,
                                                  output_cost))

    def asw_hCostAggregation(self, W: int, H: int, params: AswParams, supp_left: int, supp_right: int, vertical_cost: int,
                             denom_v: int | None, output_cost: int):
        p = params.c()
        self._check(self.lib.asw_hCostAggregation(self.h, W, H, C.byref(p), supp_left, supp_right, vertical_cost, denom_v,
                                                  output_cost))

    def asw_WTA(self, W: int, H: int, params: AswParams, cost: int, output: int | None, d_est_reference: int | None,
                d_est_target: int | None, output_target: int | None, confidence_reference: int | None,
                confidence_target: int | None):
        p = params.c()
        self._check(self.lib.asw_WTA(self.h, W, H, C.byref(p), cost, output, d_est_reference, d_est_target, output_target,
                                     confidence_reference, confidence_target))


class AswMulti:
    """One frame on several GPUs of this process (asw_multi_*): row bands, neighbour rows pulled over NVLink between
    iterations.  Host buffers in and out, bit-identical to AswContext.disparity on one GPU."""

    def __init__(self, devices):
        self.lib = load_library()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        st = self.lib.asw_multi_create(C.byref(h), devs, len(devices))
        if st != ASW_OK:
            raise AswError(st, f"asw_multi_create({list(devices)}) failed: {self.lib.asw_strerror(st).decode()}")
        self.h, self.devices = h, list(devices)

    def close(self):
        if getattr(self, "h", None):
            self.lib.asw_multi_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def disparity(self, left: np.ndarray, right: np.ndarray, params: AswParams | None = None, want_conf: bool = True) -> dict:
        params = params or AswParams()
        left, right = _rgba(left), _rgba(right)
        H, W, _ = left.shape
        out = {"disp_rgba": np.empty((H, W, 4), np.uint8),
               "disp_d": np.empty((H, W), np.uint8) if params.ndisp <= 256 else None,
               "conf": np.empty((H, W), np.float32) if want_conf else None}
        tm, p = CMultiTiming(), params.c()
        st = self.lib.asw_multi_disparity(self.h, left.ctypes.data, right.ctypes.data, W, H, C.byref(p), _host_ptr(out["disp_rgba"]),
                                          _host_ptr(out["disp_d"]), _host_ptr(out["conf"]), C.byref(tm))
        if st != ASW_OK:
            raise AswError(st, f"{self.lib.asw_strerror(st).decode()}: {self.lib.asw_multi_last_error(self.h).decode()}")
        out["timing"] = tm.as_dict()
        return out
