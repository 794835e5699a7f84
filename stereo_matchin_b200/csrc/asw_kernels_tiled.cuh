// asw_kernels_tiled.cuh -- the sm_100a kernel family of the fused hot path.
//
// Data layout in HBM (internal to the fused path; the per-operator ABI keeps the
// reference layouts and uses asw_kernels_basic.cuh):
//   cost / denominator volumes : [y][x][Dp]   float32, disparity innermost, Dp = D rounded
//                                up to 32 (padding planes hold finite junk, never read by WTA)
//   support tables             : [y][tap][x]  float32
// With d innermost a warp's 32 lanes are 32 consecutive disparities of one pixel, so
//   * the left-image weight wL[x,y,tap] is warp-uniform  -> one broadcast LDS.128 feeds 4 taps-ops
//   * the right-image weight wR[x-d,y,tap] is reused along the (x+1,d+1) diagonal in registers
//   * volume reads/writes are 128-byte coalesced rows.
//
// Per tap the reference does ww = wL*wR; num += ww*c; den += ww (asw_vcost_aggregation.cl:37-39).
// den does not depend on the cost volume, so it is computed once (iteration 0, FIRST=true,
// with exactly the reference's accumulation order) and re-read afterwards: bit-identical
// results for 2 instead of 3 FP32 instructions per tap in iterations 1..r-1.
//
// Kernels are specialised for the reference's window (radius 16, 33 taps).
#pragma once
#include <stdlib.h>

#include "asw_common.cuh"

namespace asw {

constexpr int kR = 16;
constexpr int kT = 2 * kR + 1;

__host__ __device__ inline int padded_D(int D) { return (D + 31) & ~31; }

// ---------------------------------------------------------------------------------------
// raw cost, kernels/asw_aggr.cl:3-23, into [y][x][Dp].  One warp per pixel, lanes = d.
__global__ void k_raw_t(const uint32_t* __restrict__ L, const uint32_t* __restrict__ R, Band b, int ylo, int yhi, int D,
                        int Dp, float trunc, float* __restrict__ cost) {
    const int x = blockIdx.x * blockDim.y + threadIdx.y;
    const int y = ylo + blockIdx.y;
    if (x >= b.W || y >= yhi) return;
    const uint32_t lp = L[(size_t)y * b.W + x];
    const uint32_t* rrow = R + (size_t)y * b.W;
    float* o = cost + ((size_t)(y - b.y_off) * b.W + x) * Dp;
    for (int d = threadIdx.x; d < Dp; d += 32) {
        float v = 0.0f;
        if (d < D) v = fminf(sad_rgb(lp, rrow[max(x - d, 0)]), trunc);
        o[d] = v;
    }
}

// support tables, kernels/asw_vsupport.cl:3-27 / asw_hsupport.cl:3-28, into [y][tap][x]
template <bool VERTICAL>
__global__ void k_support_t(const uint32_t* __restrict__ img, Band b, int ylo, int yhi, float gamma_c, float gamma_p,
                            float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = ylo + blockIdx.y;
    const int i = blockIdx.z;
    if (x >= b.W || y >= yhi) return;
    int qx = x, qy = y;
    if (VERTICAL) qy = clampi(y + i - kR, 0, b.H - 1); else qx = clampi(x + i - kR, 0, b.W - 1);
    float sad = sad_rgb(img[(size_t)y * b.W + x], img[(size_t)qy * b.W + qx]);
    float c_diff = __fdiv_rn(-sad, gamma_c);
    float g_dist = __fdiv_rn((float)abs(VERTICAL ? y - qy : x - qx), gamma_p);
    out[((size_t)(y - b.y_off) * kT + i) * b.W + x] = (float)exp((double)__fsub_rn(c_diff, g_dist));
}

// ---------------------------------------------------------------------------------------
// Vertical pass, kernels/asw_vcost_aggregation.cl:11-44.
// Thread = diagonal run of NJ outputs (x0+j, e+j), j<NJ, times a run of NY output rows.
// All NJ outputs of a row share ONE right weight wR[x0-e, y, tap] (x-d is constant on the
// diagonal); the left weights wL[x0+j, y, tap] are warp-uniform (lanes differ in e only).
// Input rows are streamed: row yy is loaded once and feeds tap (yy - y + 16) of every output
// row y of the run, which visits each output's taps in the reference's order 0..32.
template <int NJ, int NY, bool FIRST>
__global__ void __launch_bounds__(256) k_vagg_t(const float* __restrict__ wL, const float* __restrict__ wR,
                                                const float* __restrict__ cin, float* __restrict__ den_vol, Band b,
                                                int ylo, int yhi, int Dp, float* __restrict__ cout) {
    static_assert(NJ == 4, "left weights are fetched as one float4 per tap");
    constexpr int XW = 8 * NJ;  // columns per CTA (8 warps = 8 x-tiles)
    extern __shared__ float sWL[];  // [NY][kT][XW]
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int W = b.W;
    const int xg = blockIdx.x * XW, x0 = xg + NJ * w;
    const int y0 = ylo + blockIdx.y * NY;
    const int kmax = min(NY, yhi - y0) - 1;  // last valid output row of the run (uniform)
    const int e = blockIdx.z * 32 + lane - (NJ - 1);

    for (int idx = tid; idx < NY * kT * XW; idx += 256) {
        const int xx = idx % XW, ki = idx / XW;
        const int k = ki / kT, i = ki - k * kT;
        const int yl = y0 + min(k, kmax) - b.y_off;
        sWL[idx] = wL[((size_t)yl * kT + i) * W + min(xg + xx, W - 1)];
    }
    __syncthreads();

    const int col = clampi(x0 - e, 0, W - 1);  // max(x-d,0) of every output on the diagonal
    size_t coff[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) coff[j] = (size_t)min(x0 + j, W - 1) * Dp + clampi(e + j, 0, Dp - 1);

    float acc[NY][NJ], den[FIRST ? NY : 1][NJ];
#pragma unroll
    for (int k = 0; k < NY; k++)
#pragma unroll
        for (int j = 0; j < NJ; j++) { acc[k][j] = 0.00001f; if (FIRST) den[k][j] = 0.00001f; }

    const float* wRp = wR + (size_t)(y0 - b.y_off) * kT * W + col;
    const size_t wr_row = (size_t)kT * W;
    const size_t crow = (size_t)W * Dp;

    // All global loads of a step are issued up front (tap indices clamped so every address is
    // valid), the next step's loads are in flight while this step's FMAs run.
    float cj[NJ], wr[NY];
    auto load_step = [&](int s, float (&c)[NJ], float (&r)[NY]) {
        const int yy = clampi(clampi(y0 - kR + s, 0, b.H - 1) - b.y_off, 0, b.Hb - 1);
#pragma unroll
        for (int j = 0; j < NJ; j++) c[j] = __ldg(cin + (size_t)yy * crow + coff[j]);
#pragma unroll
        for (int k = 0; k < NY; k++)
            r[k] = __ldg(wRp + (size_t)min(k, kmax) * wr_row + (size_t)clampi(s - k, 0, kT - 1) * W);
    };
    load_step(0, cj, wr);
#pragma unroll 1
    for (int s = 0; s < NY + 2 * kR; s++) {
        float cn[NJ], wn[NY];
        load_step(min(s + 1, NY + 2 * kR - 1), cn, wn);
#pragma unroll
        for (int k = 0; k < NY; k++) {
            const int i = s - k;
            if (i >= 0 && i < kT && k <= kmax) {
                const float4 wl = *reinterpret_cast<const float4*>(&sWL[(k * kT + i) * XW + NJ * w]);
                const float wlv[4] = {wl.x, wl.y, wl.z, wl.w};
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const float ww = __fmul_rn(wlv[j], wr[k]);
                    acc[k][j] = __fmaf_rn(ww, cj[j], acc[k][j]);
                    if (FIRST) den[k][j] = __fadd_rn(den[k][j], ww);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; j++) cj[j] = cn[j];
#pragma unroll
        for (int k = 0; k < NY; k++) wr[k] = wn[k];
    }

#pragma unroll
    for (int k = 0; k < NY; k++) {
        if (k <= kmax) {
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int d = e + j;
                if (d >= 0 && d < Dp && x0 + j < W) {
                    const size_t o = ((size_t)(y0 + k - b.y_off) * W + x0 + j) * Dp + d;
                    float dn;
                    if (FIRST) { dn = den[k][j]; den_vol[o] = dn; } else dn = __ldg(den_vol + o);
                    cout[o] = __fdiv_rn(acc[k][j], dn);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Horizontal pass, kernels/asw_hcost_aggregation.cl:12-44.
// CTA = one image row x 64 columns, looping over chunks of 4*DL disparities.
// Thread = 8 consecutive x (x-run) x 4 consecutive d (d-run); lanes of a warp are DL d-runs.
// Taps are fully unrolled; the 8x4 window of input costs lives in registers and slides by
// one column per tap (static register renaming), the 8 left weights are a broadcast load,
// the 11 right weights wR[x-d] cover the (x,d) rectangle's diagonals.
template <int DL, bool FIRST>
__global__ void __launch_bounds__(8 * DL) k_hagg_t(const float* __restrict__ wL, const float* __restrict__ wR,
                                                   const float* __restrict__ cin, float* __restrict__ den_vol, Band b,
                                                   int ylo, int Dp, float* __restrict__ cout) {
    constexpr int TX = 64, DC = 4 * DL, NCOL = TX + DC, NT = 8 * DL, XH = TX + 2 * kR;
    extern __shared__ float4 smem4[];
    float4* sC4 = smem4;                                       // [XH][DC/4]
    float* sWL = reinterpret_cast<float*>(sC4 + XH * (DC / 4));  // [kT][TX]
    float* sWR = sWL + kT * TX;                                  // [kT][NCOL]
    const int tid = threadIdx.x, dl = tid % DL, xr = tid / DL;
    const int W = b.W;
    const int yl = ylo + blockIdx.y - b.y_off;
    const int x0 = blockIdx.x * TX;
    const float* wLrow = wL + (size_t)yl * kT * W;
    const float* wRrow = wR + (size_t)yl * kT * W;
    const float* crow = cin + (size_t)yl * W * Dp;

    for (int idx = tid; idx < kT * TX; idx += NT) {
        const int i = idx / TX, xx = idx - i * TX;
        sWL[idx] = wLrow[(size_t)i * W + min(x0 + xx, W - 1)];
    }

    for (int dc = 0; dc < Dp; dc += DC) {
        const int colbase = x0 - dc - DC;
        for (int idx = tid; idx < kT * NCOL; idx += NT) {
            const int i = idx / NCOL, cc = idx - i * NCOL;
            sWR[idx] = wRrow[(size_t)i * W + clampi(colbase + cc, 0, W - 1)];
        }
        for (int idx = tid; idx < XH * (DC / 4); idx += NT) {
            const int xx = idx / (DC / 4), q = idx - xx * (DC / 4);
            const int xs = clampi(x0 - kR + xx, 0, W - 1), d4 = dc + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d4 < Dp) v = *reinterpret_cast<const float4*>(crow + (size_t)xs * Dp + d4);
            sC4[idx] = v;
        }
        __syncthreads();

        if (dc + 4 * dl < Dp) {
            float acc[8][4], den[FIRST ? 8 : 1][4];
            float4 win[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
#pragma unroll
                for (int m = 0; m < 4; m++) { acc[j][m] = 0.00001f; if (FIRST) den[j][m] = 0.00001f; }
                win[j] = sC4[(8 * xr + j) * (DC / 4) + dl];
            }
            const int cc0 = 8 * xr - 4 * dl + DC - 4;
#pragma unroll
            for (int i = 0; i < kT; i++) {
                const float4 la = *reinterpret_cast<const float4*>(&sWL[i * TX + 8 * xr]);
                const float4 lb = *reinterpret_cast<const float4*>(&sWL[i * TX + 8 * xr + 4]);
                const float4 r0 = *reinterpret_cast<const float4*>(&sWR[i * NCOL + cc0]);
                const float4 r1 = *reinterpret_cast<const float4*>(&sWR[i * NCOL + cc0 + 4]);
                const float4 r2 = *reinterpret_cast<const float4*>(&sWR[i * NCOL + cc0 + 8]);
                const float wl[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
                const float wr[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 c4 = win[(j + i) & 7];
                    const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        const float ww = __fmul_rn(wl[j], wr[j - m + 4]);   // wR[x0+8xr+j - (dc+4dl+m)]
                        acc[j][m] = __fmaf_rn(ww, cv[m], acc[j][m]);
                        if (FIRST) den[j][m] = __fadd_rn(den[j][m], ww);
                    }
                }
                if (i + 1 < kT) win[i & 7] = sC4[(8 * xr + 8 + i) * (DC / 4) + dl];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int x = x0 + 8 * xr + j;
                if (x < W) {
                    const size_t o = ((size_t)yl * W + x) * Dp + dc + 4 * dl;
                    float4 dn;
                    if (FIRST) {
                        dn = make_float4(den[j][0], den[j][1], den[j][2], den[j][3]);
                        *reinterpret_cast<float4*>(den_vol + o) = dn;
                    } else {
                        dn = *reinterpret_cast<const float4*>(den_vol + o);
                    }
                    float4 r;
                    r.x = __fdiv_rn(acc[j][0], dn.x);
                    r.y = __fdiv_rn(acc[j][1], dn.y);
                    r.z = __fdiv_rn(acc[j][2], dn.z);
                    r.w = __fdiv_rn(acc[j][3], dn.w);
                    *reinterpret_cast<float4*>(cout + o) = r;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// WTA left part, kernels/asw_wta.cl:25-47,70,73,76-77, on the [y][x][Dp] volume.
// One warp per pixel: each lane scans d = lane, lane+32, ... (ascending), then the 32
// partial (min1, min2, argmin) triples are merged with warp shuffles.
__global__ void k_wta_t(const float* __restrict__ cost, Band b, int ylo, int yhi, int D, int Dp, int out_y0,
                        uint32_t* __restrict__ out_rgba, uint8_t* __restrict__ out_d, float* __restrict__ conf) {
    const int x = blockIdx.x * blockDim.y + threadIdx.y;
    const int y = ylo + blockIdx.y;
    if (x >= b.W || y >= yhi) return;
    const float* c = cost + ((size_t)(y - b.y_off) * b.W + x) * Dp;
    Min2 m;
    m.init();
    for (int d = threadIdx.x; d < D; d += 32) m.push(c[d], d);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float oc = __shfl_xor_sync(0xffffffffu, m.cur, off);
        const float ol = __shfl_xor_sync(0xffffffffu, m.last, off);
        const int oa = __shfl_xor_sync(0xffffffffu, m.arg, off);
        m.merge(oc, ol, oa);
    }
    if (threadIdx.x == 0) {
        const size_t o = (size_t)(y - out_y0) * b.W + x;
        if (out_rgba) {
            const uint32_t v = D > 1 ? q8(__fdiv_rn((float)m.arg, (float)(D - 1))) : 0u;
            out_rgba[o] = v | (v << 8) | (v << 16) | 0xff000000u;
        }
        if (out_d) out_d[o] = (uint8_t)m.arg;
        if (conf) conf[o] = __fdiv_rn(__fsub_rn(m.last, m.cur), m.last);
    }
}

// Convert the internal [y][x][Dp] volume to the reference layout x + W*y + W*H*d (for
// asw_final_volume consumers and the parity tests of aggregated costs).
__global__ void k_volume_to_ref(const float* __restrict__ vol, Band b, int ylo, int yhi, int D, int Dp, int out_y0,
                                int out_rows, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int y = ylo + blockIdx.z;
    const int x0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    if (y >= yhi) return;
    {
        const int x = x0 + threadIdx.y, d = d0 + threadIdx.x;
        tile[threadIdx.y][threadIdx.x] = (x < b.W && d < Dp) ? vol[((size_t)(y - b.y_off) * b.W + x) * Dp + d] : 0.f;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, d = d0 + threadIdx.y;
    if (x < b.W && d < D) out[((size_t)d * out_rows + (y - out_y0)) * b.W + x] = tile[threadIdx.x][threadIdx.y];
}

// ---------------------------------------------------------------------------------------
// host-side launchers

struct TiledCfg {
    int v_ny = 8;  // output rows per thread in the vertical pass (8 or 16)
};

inline TiledCfg& tiled_cfg() {
    static TiledCfg c = [] {
        TiledCfg t;
        if (const char* e = getenv("ASW_V_NY")) t.v_ny = atoi(e) == 16 ? 16 : 8;   // tuning knob for experiments
        return t;
    }();
    return c;
}

template <typename K>
inline cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

constexpr size_t vagg_smem(int NY) { return sizeof(float) * NY * kT * 32; }
constexpr size_t hagg_smem(int DL) { return sizeof(float) * ((64 + 2 * kR) * 4 * DL + kT * 64 + kT * (64 + 4 * DL)); }

inline cudaError_t tiled_configure() {
    cudaError_t e;
    if ((e = set_smem(k_vagg_t<4, 16, false>, vagg_smem(16)))) return e;
    if ((e = set_smem(k_vagg_t<4, 8, false>, vagg_smem(8)))) return e;
    if ((e = set_smem(k_vagg_t<4, 8, true>, vagg_smem(8)))) return e;
    if ((e = set_smem(k_hagg_t<32, false>, hagg_smem(32)))) return e;
    if ((e = set_smem(k_hagg_t<32, true>, hagg_smem(32)))) return e;
    if ((e = set_smem(k_hagg_t<16, false>, hagg_smem(16)))) return e;
    if ((e = set_smem(k_hagg_t<16, true>, hagg_smem(16)))) return e;
    if ((e = set_smem(k_hagg_t<8, false>, hagg_smem(8)))) return e;
    if ((e = set_smem(k_hagg_t<8, true>, hagg_smem(8)))) return e;
    return cudaSuccess;
}

inline bool tiled_supported(int radius) { return radius == kR; }

inline cudaError_t launch_raw_t(cudaStream_t st, const uint8_t* l, const uint8_t* r, Band b, int ylo, int yhi, int D, float trunc,
                                float* cost) {
    if (yhi <= ylo) return cudaSuccess;
    dim3 blk(32, 8), grd((b.W + 7) / 8, yhi - ylo);
    k_raw_t<<<grd, blk, 0, st>>>((const uint32_t*)l, (const uint32_t*)r, b, ylo, yhi, D, padded_D(D), trunc, cost);
    return cudaGetLastError();
}

inline cudaError_t launch_support_t(cudaStream_t st, bool vertical, const uint8_t* img, Band b, int ylo, int yhi, float gc,
                                    float gp, float* out) {
    if (yhi <= ylo) return cudaSuccess;
    dim3 grd((b.W + 127) / 128, yhi - ylo, kT);
    if (vertical) k_support_t<true><<<grd, 128, 0, st>>>((const uint32_t*)img, b, ylo, yhi, gc, gp, out);
    else k_support_t<false><<<grd, 128, 0, st>>>((const uint32_t*)img, b, ylo, yhi, gc, gp, out);
    return cudaGetLastError();
}

inline cudaError_t launch_vagg_t(cudaStream_t st, bool first, Band b, int ylo, int yhi, int D, const float* wL, const float* wR,
                                 const float* cin, float* den, float* cout) {
    if (yhi <= ylo) return cudaSuccess;
    const int Dp = padded_D(D);
    const int ez = (Dp + 3 + 31) / 32;  // e runs over [-(NJ-1), Dp-1]
    if (first) {
        dim3 grd((b.W + 31) / 32, (yhi - ylo + 7) / 8, ez);
        k_vagg_t<4, 8, true><<<grd, 256, vagg_smem(8), st>>>(wL, wR, cin, den, b, ylo, yhi, Dp, cout);
    } else if (tiled_cfg().v_ny == 16) {
        dim3 grd((b.W + 31) / 32, (yhi - ylo + 15) / 16, ez);
        k_vagg_t<4, 16, false><<<grd, 256, vagg_smem(16), st>>>(wL, wR, cin, den, b, ylo, yhi, Dp, cout);
    } else {
        dim3 grd((b.W + 31) / 32, (yhi - ylo + 7) / 8, ez);
        k_vagg_t<4, 8, false><<<grd, 256, vagg_smem(8), st>>>(wL, wR, cin, den, b, ylo, yhi, Dp, cout);
    }
    return cudaGetLastError();
}

inline cudaError_t launch_hagg_t(cudaStream_t st, bool first, Band b, int ylo, int yhi, int D, const float* wL, const float* wR,
                                 const float* cin, float* den, float* cout) {
    if (yhi <= ylo) return cudaSuccess;
    const int Dp = padded_D(D);
    dim3 grd((b.W + 63) / 64, yhi - ylo);
#define ASW_H_LAUNCH(DLV)                                                                                          \
    do {                                                                                                           \
        if (first) k_hagg_t<DLV, true><<<grd, 8 * DLV, hagg_smem(DLV), st>>>(wL, wR, cin, den, b, ylo, Dp, cout);   \
        else k_hagg_t<DLV, false><<<grd, 8 * DLV, hagg_smem(DLV), st>>>(wL, wR, cin, den, b, ylo, Dp, cout);        \
    } while (0)
    if (Dp <= 32) ASW_H_LAUNCH(8);
    else if (Dp <= 64) ASW_H_LAUNCH(16);
    else ASW_H_LAUNCH(32);
#undef ASW_H_LAUNCH
    return cudaGetLastError();
}

inline cudaError_t launch_wta_t(cudaStream_t st, Band b, int ylo, int yhi, int out_y0, int D, const float* cost, uint8_t* rgba,
                                uint8_t* dd, float* conf) {
    if (yhi <= ylo) return cudaSuccess;
    dim3 blk(32, 8), grd((b.W + 7) / 8, yhi - ylo);
    k_wta_t<<<grd, blk, 0, st>>>(cost, b, ylo, yhi, D, padded_D(D), out_y0, (uint32_t*)rgba, dd, conf);
    return cudaGetLastError();
}

inline cudaError_t launch_volume_to_ref(cudaStream_t st, Band b, int ylo, int yhi, int D, const float* vol, float* out) {
    if (yhi <= ylo) return cudaSuccess;
    const int Dp = padded_D(D);
    dim3 blk(32, 32), grd((b.W + 31) / 32, (Dp + 31) / 32, yhi - ylo);
    k_volume_to_ref<<<grd, blk, 0, st>>>(vol, b, ylo, yhi, D, Dp, ylo, yhi - ylo, out);
    return cudaGetLastError();
}

}  // namespace asw
