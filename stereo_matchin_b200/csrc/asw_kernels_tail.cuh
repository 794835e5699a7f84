// asw_kernels_tail.cuh -- the consumers of the hot path ("next" rows of SURVEY.md section 8f):
// L/R consistency check, iterative disparity refinement, penalised WTA, final 3x3 median.
// One thread per pixel, reference buffer layouts; every quirk of the reference kernels that its
// committed PNGs depend on is kept (marked QUIRK).  Arithmetic order and FMA use match
// oracle/asw_tail_oracle.c (use_fma = 1) so the comparison with the oracle is bit-exact.
#pragma once
#include "asw_common.cuh"

namespace asw {

// kernels/consist.cl:3-34 `Constistency(ref, tar, confidence_ref, confidence_tar, output, output_red)`
__global__ void k_consistency(const uint32_t* __restrict__ ref, const uint32_t* __restrict__ tar, int n, float dscale,
                              float* __restrict__ conf_ref, float* __restrict__ conf_tar, uint32_t* __restrict__ out,
                              uint32_t* __restrict__ out_red) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t r = ref[p], t = tar[p];
    uint32_t o = 0, ored = 0;
    bool ok0 = false;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float pr = __fmul_rn(__fdiv_rn((float)((r >> (8 * c)) & 0xff), 255.0f), dscale);      // consist.cl:20
        const float pt = __fmul_rn(__fdiv_rn((float)((t >> (8 * c)) & 0xff), 255.0f), dscale);      // :21
        const bool ok = fabsf(__fsub_rn(pt, pr)) < 1.001f;                                          // :25
        if (c == 0) ok0 = ok;
        o |= q8(__fdiv_rn(ok ? pr : pt, dscale)) << (8 * c);                                        // :30,32
        const float red = (c == 0 || c == 3) ? 1.0f : 0.0f;                                         // :23
        ored |= q8(ok ? __fdiv_rn(pr, dscale) : red) << (8 * c);                                    // :25,33
    }
    if (out) out[p] = o;
    if (out_red) out_red[p] = ored;
    if (conf_ref && !ok0) conf_ref[p] = 0.0f;                                                       // :27
    if (conf_tar && !ok0) conf_tar[p] = 0.0f;                                                       // :28
}

// supp_v / supp_h, kernels/asw_refinement_v.cl:2-10: exp(-SAD/10.94 - dist/118.78)
__device__ __forceinline__ float supp_ref(uint32_t p, uint32_t q, int dist) {
    const float c_diff = __fdiv_rn(-sad_rgb(p, q), 10.94f);
    const float g_dist = __fdiv_rn((float)dist, 118.78f);
    return (float)exp((double)__fsub_rn(c_diff, g_dist));
}

// kernels/asw_refinement_v.cl:13-51 `asw_ref_v(input, input_est, confidence, output_REF)`
__global__ void k_ref_v(const uint32_t* __restrict__ img, const uint32_t* __restrict__ est, const float* __restrict__ conf, int W,
                        int H, int R, float dscale, float* __restrict__ out) {
    const int nxb = (W + blockDim.x - 1) / blockDim.x;   // flat grid: block = (row, 128-column strip)
    const int x = (blockIdx.x % nxb) * blockDim.x + threadIdx.x, y = blockIdx.x / nxb;
    if (x >= W) return;
    const size_t n = (size_t)W * H;
    const uint32_t p = img[(size_t)y * W + x];
    float num = 0.00001f, den = 0.00001f;
    for (int i = 0; i < 2 * R + 1; i++) {
        const int yy = clampi(y + i - R, 0, H - 1);
        const size_t qi = (size_t)yy * W + x;
        const float Dx = __fmul_rn(__fdiv_rn((float)(est[qi] & 0xff), 255.0f), dscale);   // :38
        const float ww = supp_ref(p, img[qi], abs(y - yy));                                 // :39
        const float F = conf[qi];                                                           // :40
        num = __fmaf_rn(__fmul_rn(ww, F), Dx, num);                                         // :42
        den = __fmaf_rn(ww, F, den);                                                        // :43
    }
    out[(size_t)y * W + x] = __fdiv_rn(num, den);
    out[(size_t)y * W + x + n] = den;
}

// kernels/asw_refinement_h.cl:16-53 `asw_ref_h(input, confidence, input_REF, output_REF)`
__global__ void k_ref_h(const uint32_t* __restrict__ img, const float* __restrict__ conf, const float* __restrict__ in, int W, int H,
                        int R, float* __restrict__ out) {
    const int nxb = (W + blockDim.x - 1) / blockDim.x;   // flat grid: block = (row, 128-column strip)
    const int x = (blockIdx.x % nxb) * blockDim.x + threadIdx.x, y = blockIdx.x / nxb;
    if (x >= W) return;
    const size_t n = (size_t)W * H;
    const uint32_t p = img[(size_t)y * W + x];
    float num = 0.00001f, den = 0.00001f;
    for (int i = 0; i < 2 * R + 1; i++) {
        const int xx = clampi(x + i - R, 0, W - 1);
        const size_t qi = (size_t)y * W + xx;
        const float ww = supp_ref(p, img[qi], abs(x - xx));                                 // :41
        const float wf = __fmul_rn(ww, conf[qi]);                                           // :42
        num = __fmaf_rn(__fmul_rn(wf, in[qi]), in[qi + n], num);                            // :44
        den = __fmaf_rn(wf, in[qi + n], den);                                               // :45
    }
    out[(size_t)y * W + x] = __fdiv_rn(num, den);
    out[(size_t)y * W + x + n] = den;
}

// kernels/asw_wta_ref.cl:2-68 `asw_WTA_REF(agg_d, ref, ref_target, output, output_target, disp_ref,
// disp_ref_target, confidence, confidence_target)`
__global__ void k_wta_ref(const float* __restrict__ cost, const float* __restrict__ ref, const float* __restrict__ ref_t, int W, int H,
                          int D, uint32_t* __restrict__ out, uint32_t* __restrict__ out_t, float* __restrict__ disp_ref,
                          float* __restrict__ disp_ref_t, float* __restrict__ confidence) {
    const int nxb = (W + blockDim.x - 1) / blockDim.x;   // flat grid: block = (row, 128-column strip)
    const int x = (blockIdx.x % nxb) * blockDim.x + threadIdx.x, y = blockIdx.x / nxb;
    if (x >= W) return;
    const size_t n = (size_t)W * H, p = (size_t)y * W + x;
    Min2 m;
    m.init();
    const float a = __fmul_rn(0.085f, ref[p + n]), r0 = ref[p];                             // :26
    for (int i = 0; i < D; i++) m.push(__fmaf_rn(a, fabsf(__fsub_rn(r0, (float)i)), cost[p + n * i]), i);
    Min2 t;
    t.init();
    t.arg = m.arg;
    const float at = __fmul_rn(0.085f, ref_t[p + n]), rt0 = ref_t[p];
    for (int i = 0; i < m.arg; i++) {
        const int xq = max(0, x - i), b = xq - x + m.arg;
        // QUIRK asw_wta_ref.cl:46: the penalty uses the loop index i, not the sampled disparity b
        t.push(__fmaf_rn(at, fabsf(__fsub_rn(rt0, (float)i)), cost[(size_t)xq + (size_t)W * y + n * b]), b);
    }
    const float scale = (float)(D - 1);
    const uint32_t v = D > 1 ? q8(__fdiv_rn((float)m.arg, scale)) : 0u, vt = D > 1 ? q8(__fdiv_rn((float)t.arg, scale)) : 0u;
    out[p] = v | (v << 8) | (v << 16) | 0xff000000u;
    out_t[p] = vt | (vt << 8) | (vt << 16) | 0xff000000u;
    if (disp_ref) disp_ref[p] = (float)m.arg;
    if (disp_ref_t) disp_ref_t[p] = (float)t.arg;
    // QUIRK :63,66: both confidences go to the SAME buffer, the target one last; confidence_target
    // is never rewritten by this kernel
    confidence[p] = __fdiv_rn(__fsub_rn(t.last, t.cur), t.last);
}

// kernels/median.cl:58-88 `Median(input, output)`: per-channel 3x3 median, clamp-to-edge
__global__ void k_median(const uint32_t* __restrict__ in, int W, int H, uint32_t* __restrict__ out) {
    const int nxb = (W + blockDim.x - 1) / blockDim.x;   // flat grid: block = (row, 128-column strip)
    const int x = (blockIdx.x % nxb) * blockDim.x + threadIdx.x, y = blockIdx.x / nxb;
    if (x >= W) return;
    uint32_t s[9];
    int k = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) s[k++] = in[(size_t)clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)];
    uint32_t o = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t v[9];
#pragma unroll
        for (int i = 0; i < 9; i++) v[i] = (s[i] >> (8 * c)) & 0xff;
        // the min/max exchange network of median.cl:82-85 (McGuire): v[4] ends up as the median
#define S2(a, b) { const uint32_t lo = min(v[a], v[b]); v[b] = max(v[a], v[b]); v[a] = lo; }
        S2(0, 3) S2(1, 4) S2(2, 5) S2(0, 1) S2(0, 2) S2(4, 5) S2(3, 5)          // mnmx6(0..5)
        S2(1, 2) S2(3, 4) S2(1, 3) S2(1, 6) S2(4, 6) S2(2, 6)                   // mnmx5(1,2,3,4,6)
        S2(2, 3) S2(4, 7) S2(2, 4) S2(3, 7)                                     // mnmx4(2,3,4,7)
        S2(4, 8) S2(3, 8) S2(3, 4)                                              // mnmx3(3,4,8)
#undef S2
        o |= v[4] << (8 * c);
    }
    out[(size_t)y * W + x] = o;
}

}  // namespace asw
