// asw_common.cuh -- shared device helpers and the band geometry used by every kernel.
//
// Numerics contract (matches oracle/asw_oracle.c, which restates the reference kernels):
// every float operation is written with an explicit round-to-nearest intrinsic so that
// nvcc never contracts or reorders anything the reference's arithmetic does not allow;
// the only fused operation is the tap accumulation num = fma(ww, c, num), which the
// reference's OpenCL build (no options, FP_CONTRACT on, main.cpp:211) permits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace asw {

constexpr int kR = 16;             // the reference's window radius (asw_vsupport.cl:19): the fused kernels are specialised for it
constexpr int kT = 2 * kR + 1;     // 33 taps

template <typename K>
inline cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// A row band of a W x H frame held in band-local buffers: global row y lives at local row
// y - y_off, the buffers have Hb rows.  For a whole frame y_off = 0, Hb = H.
struct Band {
    int W, H;      // full frame size (clamp-to-edge applies at these borders only)
    int y_off;     // first global row held in the band buffers
    int Hb;        // rows held
    __host__ __device__ size_t plane() const { return (size_t)W * (size_t)Hb; }
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// read_imagef(CL_UNORM_INT8) * 255 (asw_aggr.cl:12): v/255.0f then *255, two roundings.
__device__ __forceinline__ float px(uint32_t v) { return __fmul_rn(__fdiv_rn((float)v, 255.0f), 255.0f); }

// |a.r-b.r| + |a.g-b.g| + |a.b-b.b| in the reference's left-to-right order (asw_aggr.cl:19)
__device__ __forceinline__ float sad_rgb(uint32_t a, uint32_t b) {
    float d0 = fabsf(__fsub_rn(px(a & 0xff), px(b & 0xff)));
    float d1 = fabsf(__fsub_rn(px((a >> 8) & 0xff), px((b >> 8) & 0xff)));
    float d2 = fabsf(__fsub_rn(px((a >> 16) & 0xff), px((b >> 16) & 0xff)));
    return __fadd_rn(__fadd_rn(d0, d1), d2);
}

// write_imagef(float -> UNORM8) as observed in the reference's committed PNGs:
// round-half-down of the float32 product (oracle q8()).
__device__ __forceinline__ uint32_t q8(float f) {
    float t = __fmul_rn(f, 255.0f);
    if (!(t > 0.0f)) return 0u;
    if (t >= 255.0f) return 255u;
    return (uint32_t)ceilf(__fsub_rn(t, 0.5f));
}

// Two-smallest tracker of asw_wta.cl:43-46 (strict '<', lowest index wins ties).
struct Min2 {
    float cur, last;
    int arg;
    __device__ __forceinline__ void init() { cur = 100000.0f; last = 100000.0f; arg = 0; }
    __device__ __forceinline__ void push(float t, int i) {
        last = t < last ? t : last;
        arg = t < cur ? i : arg;
        last = t < cur ? cur : last;
        cur = t < cur ? t : cur;
    }
    // Merge another tracker that scanned a disjoint index set.  The pair (cur,last) is the
    // two smallest values of the multiset union and arg the lowest index attaining the
    // minimum, i.e. exactly what one sequential scan over the union produces.
    __device__ __forceinline__ void merge(float ocur, float olast, int oarg) {
        bool take = (ocur < cur) || (ocur == cur && oarg < arg);
        float lo = take ? ocur : cur;
        float other_min = take ? cur : ocur;      // the larger of the two minima
        float second = fminf(other_min, fminf(last, olast));
        arg = take ? oarg : arg;
        cur = lo;
        last = second;
    }
};

}  // namespace asw
