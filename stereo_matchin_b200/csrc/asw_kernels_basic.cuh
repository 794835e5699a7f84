// asw_kernels_basic.cuh -- one-thread-per-output CUDA kernels with the reference's data
// layouts.  They back the per-operator C-ABI entry points (asw_Aggr, asw_vSupport, ...),
// work for any radius / ndisp / pitch, and serve as the on-device cross-check of the TMA-fed
// kernels (kernel family 1).  Threads map to x fastest so every global access is coalesced.
#pragma once
#include "asw_common.cuh"

namespace asw {

// kernels/asw_aggr.cl:3-23 -- cost[x,y,d] = min(SAD(L(x,y), R(max(x-d,0),y)), trunc)
__global__ void k_raw_cost(const uint32_t* __restrict__ L, const uint32_t* __restrict__ R, Band b, int ylo, int yhi,
                           int D, float trunc, float* __restrict__ cost) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = ylo + blockIdx.y;
    if (x >= b.W || y >= yhi) return;
    uint32_t lp = L[(size_t)y * b.W + x];
    const uint32_t* rrow = R + (size_t)y * b.W;
    float* o = cost + (size_t)(y - b.y_off) * b.W + x;
    const size_t plane = b.plane();
    for (int d = blockIdx.z; d < D; d += gridDim.z) {
        float v = sad_rgb(lp, rrow[max(x - d, 0)]);
        o[plane * d] = fminf(v, trunc);
    }
}

// kernels/asw_vsupport.cl:3-27 / asw_hsupport.cl:3-28 -- one thread per (x, y, tap)
template <bool VERTICAL>
__global__ void k_support(const uint32_t* __restrict__ img, Band b, int ylo, int yhi, int R, float gamma_c,
                          float gamma_p, float* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = ylo + blockIdx.y;
    int i = blockIdx.z;
    if (x >= b.W || y >= yhi) return;
    int qx = x, qy = y;
    if (VERTICAL) qy = clampi(y + i - R, 0, b.H - 1); else qx = clampi(x + i - R, 0, b.W - 1);
    float sad = sad_rgb(img[(size_t)y * b.W + x], img[(size_t)qy * b.W + qx]);
    float c_diff = __fdiv_rn(-sad, gamma_c);                                   // asw_vsupport.cl:22
    float g_dist = __fdiv_rn((float)abs(VERTICAL ? y - qy : x - qx), gamma_p);  // :24 (distance of an axis shift)
    float w = (float)exp((double)__fsub_rn(c_diff, g_dist));                   // :25, correctly rounded exp
    out[b.plane() * i + (size_t)(y - b.y_off) * b.W + x] = w;
}

// kernels/asw_vcost_aggregation.cl:11-44 / asw_hcost_aggregation.cl:12-44
template <bool VERTICAL>
__global__ void k_aggregate(const float* __restrict__ sL, const float* __restrict__ sR, const float* __restrict__ cin,
                            Band b, int ylo, int yhi, int R, float* __restrict__ den_out, float* __restrict__ cout) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = ylo + blockIdx.y;
    int d = blockIdx.z;
    if (x >= b.W || y >= yhi) return;
    const size_t plane = b.plane();
    const int yl = y - b.y_off;
    const size_t idx = (size_t)yl * b.W + x;
    const size_t idx_d = (size_t)yl * b.W + max(x - d, 0);
    const float* c = cin + plane * d;
    float num = 0.00001f, den = 0.00001f;
    const int T = 2 * R + 1;
    for (int i = 0; i < T; i++) {
        float ww = __fmul_rn(sL[idx + plane * i], sR[idx_d + plane * i]);
        float cv;
        if (VERTICAL) {
            int yy = clampi(clampi(y + i - R, 0, b.H - 1) - b.y_off, 0, b.Hb - 1);
            cv = c[(size_t)yy * b.W + x];
        } else {
            cv = c[(size_t)yl * b.W + clampi(x + i - R, 0, b.W - 1)];
        }
        num = __fmaf_rn(ww, cv, num);
        den = __fadd_rn(den, ww);
    }
    cout[plane * d + idx] = __fdiv_rn(num, den);
    if (den_out) den_out[plane * d + idx] = den;
}

// kernels/asw_wta.cl:12-82 (left part :25-47, right/target part :50-67)
__global__ void k_wta(const float* __restrict__ cost, Band b, int ylo, int yhi, int D, int out_y0,
                      uint32_t* __restrict__ out_rgba, uint8_t* __restrict__ out_d, float* __restrict__ d_ref,
                      float* __restrict__ d_tar, uint32_t* __restrict__ out_tar_rgba, float* __restrict__ conf_ref,
                      float* __restrict__ conf_tar) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = ylo + blockIdx.y;
    if (x >= b.W || y >= yhi) return;
    const size_t plane = b.plane();
    const size_t row = (size_t)(y - b.y_off) * b.W;
    Min2 m;
    m.init();
    for (int i = 0; i < D; i++) m.push(cost[plane * i + row + x], i);
    const size_t o = (size_t)(y - out_y0) * b.W + x;   // outputs hold rows starting at out_y0
    const float scale = (float)(D - 1);
    if (out_rgba) {
        uint32_t v = D > 1 ? q8(__fdiv_rn((float)m.arg, scale)) : 0u;
        out_rgba[o] = v | (v << 8) | (v << 16) | 0xff000000u;
    }
    if (out_d) out_d[o] = (uint8_t)m.arg;
    if (d_ref) d_ref[o] = (float)m.arg;
    if (conf_ref) conf_ref[o] = __fdiv_rn(__fsub_rn(m.last, m.cur), m.last);
    if (d_tar || out_tar_rgba || conf_tar) {
        Min2 t;
        t.init();
        t.arg = m.arg;
        const int d_r = m.arg;
        for (int i = 0; i < d_r; i++) {
            int xq = max(0, x - i);
            // bresenham((0, x-d_r), (min_d, x), xq), asw_wta.cl:3-9,57 (integer division)
            int bb = 0;
            if ((x - d_r) - x != 0) bb = (0 - d_r) / ((x - d_r) - x) * (xq - x) + d_r;
            t.push(cost[plane * bb + row + xq], bb);
        }
        if (out_tar_rgba) {
            uint32_t v = D > 1 ? q8(__fdiv_rn((float)t.arg, scale)) : 0u;
            out_tar_rgba[o] = v | (v << 8) | (v << 16) | 0xff000000u;
        }
        if (d_tar) d_tar[o] = (float)t.arg;
        if (conf_tar) conf_tar[o] = __fdiv_rn(__fsub_rn(t.last, t.cur), t.last);
    }
}

}  // namespace asw
