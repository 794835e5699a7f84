"""Runs the measurement probes of libasw_ubench.so and prints a small JSON report."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ub = C.CDLL(os.path.join(ROOT, "stereo_matchin_b200", "libasw_ubench.so"))
ub.asw_ubench_ffma_tflops.restype = C.c_double
ub.asw_ubench_lds.restype = C.c_double
ub.asw_ubench_lds.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
rep = {"ffma_tflops": ub.asw_ubench_ffma_tflops(0), "ffma2_tflops": ub.asw_ubench_ffma_tflops(1), "lds128": {}}
names = {0: "distinct", 1: "broadcast", 2: "4_addr_quarter_warp", 3: "8_addr"}
for mode in range(4):
    for f in (0, 4, 8, 16):
        ms = C.c_double()
        v = ub.asw_ubench_lds(mode, f, C.byref(ms))
        rep["lds128"][f"{names[mode]}_fma{f}"] = {"lds_per_ns_per_sm": v, "ms": ms.value}
ub.asw_ubench_mix.restype = C.c_double
ub.asw_ubench_mix.argtypes = [C.c_int, C.c_double]
rep["mix_lane_ops_per_clk_per_sm_at_1965MHz"] = {n: ub.asw_ubench_mix(m, 1965.0) for m, n in
    enumerate(["fmul+ffma_scalar_tapops(peak 64)", "fmul2+ffma2_packed_tapops(peak 64)", "fmul_only(peak 128)", "ffma_3src_only(peak 128)"])}
ub.asw_ubench_lds_rate.restype = C.c_double
ub.asw_ubench_lds_rate.argtypes = [C.c_int, C.c_int]
pat = ["distinct", "broadcast", "l==l+16 (16 distinct)", "pairs (16 distinct)", "l mod 8 (8 distinct)", "groups of 4 (8 distinct)",
       "groups of 8 (4 distinct)"]
rep["lds_sm_cycles_per_warp_instruction"] = {f"LDS.{8 * w} {pat[p]}": round(ub.asw_ubench_lds_rate(w, p), 3) for w in (4, 8, 16) for p in range(7)}
print(json.dumps(rep, indent=1))
