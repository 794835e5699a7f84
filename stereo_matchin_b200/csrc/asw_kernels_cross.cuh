// asw_kernels_cross.cuh -- the reference's second method, "Cross-Based Local Stereo Matching Using
// Orthogonal Integral Images" (host order main.cpp:258-367): cross
// construction, raw cost on UNORM colours, running sums along x / y, cross-limited box means, initial
// WTA and cross-region voting.  Reference buffer layouts (cost: x + W*y + W*H*d, cross: 4 planes of ints).
// Arithmetic and evaluation order match oracle/cross_oracle.c so the comparison is bit-exact; in
// particular the running sums are sequential per row / column (a parallel scan would round differently).
#pragma once
#include "asw_common.cuh"

namespace asw {

__device__ __forceinline__ float unorm8(uint32_t v) { return __fdiv_rn((float)v, 255.0f); }   // read_imagef, CL_UNORM_INT8
__device__ __forceinline__ uint32_t ld_clamped(const uint32_t* __restrict__ img, int W, int H, int x, int y) {
    return img[(size_t)clampi(y, 0, H - 1) * W + clampi(x, 0, W - 1)];                         // CLAMP_TO_EDGE sampler
}

// pixels the reference's 3x3 NDRange never reaches (main.cpp:193,197): written as zeros
__global__ void k_cb_zero_border(uint32_t* __restrict__ img, int W, int H, int We, int He) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < W && (x >= We || y >= He)) img[(size_t)y * W + x] = 0u;
}

// check_all of kernels/cross.cl:25-81 (arm k is tested on the pixel at distance k + 1, as there)
__device__ __forceinline__ int cb_arm(const uint32_t* __restrict__ img, int W, int H, int x, int y, int ox, int oy, int max_arm) {
    const uint32_t p = img[(size_t)y * W + x];
    const float c0 = unorm8(p & 0xff), c1 = unorm8((p >> 8) & 0xff), c2 = unorm8((p >> 16) & 0xff);
    int arm = 1;
    for (int k = 1; k <= max_arm; k++) {
        const int nx = x + (k + 1) * ox, ny = y + (k + 1) * oy;
        const uint32_t n = ld_clamped(img, W, H, nx, ny);
        const float check = (fabsf(__fsub_rn(c0, unorm8(n & 0xff))) < 0.10f ? 1.0f : 0.0f) +
                            (fabsf(__fsub_rn(c1, unorm8((n >> 8) & 0xff))) < 0.10f ? 1.0f : 0.0f) +
                            (fabsf(__fsub_rn(c2, unorm8((n >> 16) & 0xff))) < 0.10f ? 1.0f : 0.0f);     // cross.cl:5-13
        int flag = (float)(k - arm) > 1.0f ? 1 : 0;                                                    // :15
        const int cur = 3.0f <= check ? k : arm;                                                       // :16
        flag += (nx < 0) + (ny < 0) + (W <= nx) + (H <= ny);                                           // :17-20
        arm = flag ? arm : cur;                                                                        // :22
    }
    return arm;
}

// kernels/cross.cl:83-105 `Cross(input, output)`
__global__ void k_cb_cross(const uint32_t* __restrict__ img, int W, int H, int max_arm, int* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t n = (size_t)W * H, p = (size_t)y * W + x;
    out[p] = -cb_arm(img, W, H, x, y, -1, 0, max_arm);
    out[p + n] = cb_arm(img, W, H, x, y, 1, 0, max_arm);
    out[p + 2 * n] = -cb_arm(img, W, H, x, y, 0, -1, max_arm);
    out[p + 3 * n] = cb_arm(img, W, H, x, y, 0, 1, max_arm);
}

// kernels/aggregation.cl:3-23 `Aggregation(input_l, input_r, output_cost)`; one thread per (x, y, d)
__global__ void k_cb_aggregation(const uint32_t* __restrict__ L, const uint32_t* __restrict__ R, int W, int H, float* __restrict__ cost) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, d = blockIdx.z;
    if (x >= W) return;
    const uint32_t l = L[(size_t)y * W + x], r = R[(size_t)y * W + max(x - d, 0)];
    const float s = __fadd_rn(__fadd_rn(fabsf(__fsub_rn(unorm8(l & 0xff), unorm8(r & 0xff))),
                                        fabsf(__fsub_rn(unorm8((l >> 8) & 0xff), unorm8((r >> 8) & 0xff)))),
                              fabsf(__fsub_rn(unorm8((l >> 16) & 0xff), unorm8((r >> 16) & 0xff))));
    cost[(size_t)y * W + x + (size_t)W * H * d] = s;
}

// kernels/integral_h.cl:3-17 `Integral_h(cost, size)`: running sum along x, strictly left to right.
// A block owns 32 consecutive rows of the (H*D) x W matrix; 32x32 tiles go through shared memory so global
// accesses are coalesced, and thread `row` of the first warp carries that row's sum from tile to tile.
__global__ void __launch_bounds__(256) k_cb_integral_h(float* __restrict__ cost, int W, int nrows) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;               // 32 x 8
    const int r0 = blockIdx.x * 32;
    float carry = 0.0f;
    for (int x0 = 0; x0 < W; x0 += 32) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int row = r0 + ty + 8 * k, x = x0 + tx;
            tile[ty + 8 * k][tx] = (row < nrows && x < W) ? cost[(size_t)row * W + x] : 0.0f;
        }
        __syncthreads();
        if (ty == 0) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                carry = __fadd_rn(carry, tile[tx][j]);
                tile[tx][j] = carry;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int row = r0 + ty + 8 * k, x = x0 + tx;
            if (row < nrows && x < W) cost[(size_t)row * W + x] = tile[ty + 8 * k][tx];
        }
        __syncthreads();
    }
}

// kernels/integral_v.cl:3-17 `Integral_v(cost, size)`: running sum along y; one thread per (x, d) column
__global__ void k_cb_integral_v(float* __restrict__ cost, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
    if (x >= W) return;
    float* col = cost + (size_t)W * H * d + x;
    float sum = 0.0f;
    for (int i = 0; i < H; i++) {
        sum = __fadd_rn(sum, col[(size_t)i * W]);
        col[(size_t)i * W] = sum;
    }
}

// kernels/oii_hcross.cl:1-31 / oii_vcross.cl:1-32: box mean over the intersection of the left arm at x and the
// right arm at max(0, x - d); the divisor is (plus - minus) as in the reference (one less than the pixel count)
template <bool HORIZONTAL>
__global__ void k_cb_oii(const int* __restrict__ cross_l, const int* __restrict__ cross_r, const float* __restrict__ in, int W, int H,
                         float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, d = blockIdx.z;
    if (x >= W) return;
    const size_t n = (size_t)W * H, pl = (size_t)y * W + x, pr = (size_t)y * W + max(x - d, 0);
    const int o = HORIZONTAL ? 0 : 2;
    const int minus = max(cross_r[pr + o * n], cross_l[pl + o * n]);
    const int plus = min(cross_r[pr + (o + 1) * n], cross_l[pl + (o + 1) * n]);
    const float* plane = in + n * d;
    float a, b;
    if (HORIZONTAL) {
        a = plane[(size_t)y * W + min(W - 1, x + plus)];
        b = plane[(size_t)y * W + max(0, x + minus - 1)];
    } else {
        a = plane[(size_t)min(H - 1, y + plus) * W + x];
        b = plane[(size_t)max(0, y + minus - 1) * W + x];
    }
    out[pl + n * d] = __fdiv_rn(__fsub_rn(a, b), (float)(plus - minus));
}

// kernels/init_disparity.cl:1-19 `Init_disparity(cost, output)`: strict-less argmin, lowest d wins ties
__global__ void k_cb_init_disparity(const float* __restrict__ cost, int W, int H, int D, uint32_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t n = (size_t)W * H, p = (size_t)y * W + x;
    int min_d = 0;
    float min_result = cost[p];
    for (int i = 0; i < D; i++) {
        const float c = cost[p + n * i];
        if (c < min_result) { min_d = i; min_result = c; }
    }
    const uint32_t v = D > 1 ? q8(__fdiv_rn((float)min_d, (float)(D - 1))) : 0u;
    out[p] = v | (v << 8) | (v << 16) | 0xff000000u;
}

// kernels/disparity.cl:1-41 `Disparity(input, input_cross, output)`: votes of the initial disparities over the
// pixel's cross region; the bin is (int)(v/255 * (D-1)) truncated and ties go to the larger disparity, as there.
// One warp per pixel: lanes split the region's rows, the per-pixel histogram lives in shared memory.
template <int DMAX>
__global__ void __launch_bounds__(256) k_cb_disparity(const uint32_t* __restrict__ init, const int* __restrict__ cross, int W, int H, int D,
                                                      uint32_t* __restrict__ out) {
    __shared__ int tab[8][DMAX];
    const int lane = threadIdx.x, wi = threadIdx.y;
    const int x = blockIdx.x * 8 + wi, y = blockIdx.y;
    const bool live = x < W;
    for (int i = lane; i < D; i += 32) tab[wi][i] = 0;
    __syncwarp();
    const size_t n = (size_t)W * H;
    const float scale = (float)(D - 1);
    if (live) {
        const size_t p = (size_t)y * W + x;
        const int v_minus = cross[p + 2 * n], v_plus = cross[p + 3 * n];
        for (int i = v_minus; i <= v_plus; i++) {
            const size_t q = (size_t)x + (size_t)clampi(y + i, 0, H - 1) * W;
            const int h_minus = cross[q], h_plus = cross[q + n];
            for (int j = h_minus + lane; j <= h_plus; j += 32) {
                const float v = __fmul_rn(unorm8(ld_clamped(init, W, H, x + j, y + i) & 0xff), scale);   // disparity.cl:29
                atomicAdd(&tab[wi][(int)v], 1);                                                           // :30 (counts commute)
            }
        }
    }
    __syncwarp();
    if (live && lane == 0) {
        int result = 0, result_indx = 0;
        for (int i = 0; i < D; i++)
            if (!((float)tab[wi][i] < (float)result)) { result_indx = i; result = tab[wi][i]; }          // :35-36
        const uint32_t v = D > 1 ? q8((float)((double)result_indx / (double)(D - 1))) : 0u;              // :38 (60.0 is a double literal)
        out[(size_t)y * W + x] = v | (v << 8) | (v << 16) | 0xff000000u;
    }
}

}  // namespace asw
