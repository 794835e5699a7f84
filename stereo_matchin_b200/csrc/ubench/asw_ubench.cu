// asw_ubench.cu -- measurement helpers, built as a separate library (libasw_ubench.so).
// Not part of the drop-in ABI: bench.py uses asw_ubench_ffma_tflops() to calibrate the FP32
// roofline denominator on the GPU it runs on (MEASURED_PEAKS.json only holds HBM and bf16),
// and the other probes document the shared-memory / packed-FP32 behaviour the tiled kernels
// were designed around (profiles/).
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>
#include <stdio.h>

#define UB_API extern "C" __attribute__((visibility("default")))

namespace {

template <int ILP>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = __fmaf_rn(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

// packed FP32: fma.rn.f32x2 (SASS FFMA2) on 64-bit register pairs
template <int ILP>
__global__ void k_ffma2(float* out, int iters, float a, float b) {
    unsigned long long acc[ILP], va, vb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(va) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(vb) : "f"(b));
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        float x = threadIdx.x * 0.001f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x));
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(va), "l"(vb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
}

// mix: per `ILP` FFMA, one LDS.128 with the given address pattern
//   mode 0: all lanes distinct (conflict-free), 1: all lanes same address (broadcast),
//   2: four distinct addresses (one per quarter warp), 3: eight distinct (one per 4 lanes)
template <int FMA_PER_LDS>
__global__ void k_lds_mix(float* out, int iters, int mode) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int idx;
    if (mode == 0) idx = lane;
    else if (mode == 1) idx = 0;
    else if (mode == 2) idx = lane >> 3;
    else idx = lane >> 2;
    idx += warp * 32;
    float acc[FMA_PER_LDS > 0 ? FMA_PER_LDS : 1];
#pragma unroll
    for (int i = 0; i < (FMA_PER_LDS > 0 ? FMA_PER_LDS : 1); i++) acc[i] = lane + i;
    float4 s4 = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; it++) {
        float4 v;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(&buf[(idx + it * 32) & 1023]);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
        s4.x += v.x + v.w;   // keep the load alive
#pragma unroll
        for (int i = 0; i < FMA_PER_LDS; i++) acc[i] = __fmaf_rn(acc[i], 1.0001f, v.y);
    }
    float s = s4.x;
#pragma unroll
    for (int i = 0; i < (FMA_PER_LDS > 0 ? FMA_PER_LDS : 1); i++) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

// tap-loop shaped mixes: w = a*b (FMUL); acc = fma(w, c, acc), operands in distinct registers.
//   mode 0: scalar FMUL + FFMA     mode 1: packed FMUL2 + FFMA2     mode 2: FMUL only     mode 3: FFMA only (3 distinct sources)
template <int MODE>
__global__ void k_mix(float* out, int iters, const float* __restrict__ in) {
    float a[8], b[8], c[8], acc[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = in[i] + threadIdx.x; b[i] = in[8 + i]; c[i] = in[16 + i]; }
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = i;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) { float w = __fmul_rn(a[i & 7], b[(i * 3) & 7]); acc[i] = __fmaf_rn(w, c[(i * 5) & 7], acc[i]); }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long w, aa, bb, cc, ac;
                asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a[i]), "f"(a[(i + 1) & 7]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b[(i * 3) & 7]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c[i]), "f"(c[(i + 3) & 7]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(ac) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(w) : "l"(aa), "l"(bb));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ac) : "l"(w), "l"(cc));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(ac));
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = __fmul_rn(acc[i], b[i & 7]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = __fmaf_rn(a[i & 7], c[(i * 5) & 7], acc[i]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

float time_ms(void (*launch)(void*), void* arg, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch(arg);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a);
        launch(arg);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return best;
}

struct Arg {
    float* out;
    int iters, blocks, threads, mode;
};

}  // namespace

// ---- precise shared-memory issue-rate probe -----------------------------------------------------
// 16 unrolled loads per iteration at immediate offsets (1 issue slot per load), all warps of the SM
// active; reports SM cycles per warp-level load instruction measured with clock64 on the SM itself.
//   width: 4, 8, 16 bytes per lane.  pattern: 0 all lanes distinct & contiguous, 1 full broadcast,
//   2 lanes l and l+16 share (16 distinct), 3 lane pairs (2k, 2k+1) share (16 distinct),
//   4 lanes l, l+8, l+16, l+24 share (8 distinct), 5 groups of 4 consecutive lanes share (8 distinct),
//   6 groups of 8 consecutive lanes share (4 distinct)
template <int WIDTH>
__global__ void k_lds_rate(float* out, long long* cycles, int iters, int pattern) {
    __shared__ __align__(16) float buf[8 * 1024];
    for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) buf[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int idx;
    switch (pattern) {
        case 0: idx = lane; break;
        case 1: idx = 0; break;
        case 2: idx = lane & 15; break;
        case 3: idx = lane >> 1; break;
        case 4: idx = lane & 7; break;
        case 5: idx = lane >> 2; break;
        default: idx = lane >> 3; break;
    }
    const unsigned base0 = (unsigned)__cvta_generic_to_shared(buf) + (unsigned)(idx * WIDTH) + (unsigned)(warp & 3) * 4096u;
    float s = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const unsigned base = base0 + ((unsigned)(it & 1) << 13);   // the address changes every iteration: nothing can be reused
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (WIDTH == 16) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + u * 512) : "memory");
                s += v.x;
            } else if (WIDTH == 8) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + u * 512) : "memory");
                s += v.x;
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + u * 512) : "memory");
                s += v;
            }
        }
    }
    const long long t1 = clock64();
    // the CTA's span: earliest start to latest end over its warps (the scheduler favours old warps)
    __shared__ unsigned long long tmin, tmax;
    if (threadIdx.x == 0) { tmin = ~0ull; tmax = 0ull; }
    __syncthreads();
    if (lane == 0) { atomicMin(&tmin, (unsigned long long)t0); atomicMax(&tmax, (unsigned long long)t1); }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (long long)(tmax - tmin);
    if (s == 123.456f) out[0] = s;
}

// SM cycles per warp-level LDS instruction (all 32 warps of one 1024-thread CTA per SM issuing)
UB_API double asw_ubench_lds_rate(int width, int pattern) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* out;
    long long* cyc;
    cudaMalloc(&out, 64);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const int iters = 2048, threads = 1024;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (width == 16) k_lds_rate<16><<<sms, threads>>>(out, cyc, iters, pattern);
        else if (width == 8) k_lds_rate<8><<<sms, threads>>>(out, cyc, iters, pattern);
        else k_lds_rate<4><<<sms, threads>>>(out, cyc, iters, pattern);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaDeviceSynchronize();
    if (getenv("ASW_UBENCH_WALL")) {   // wall-clock view: ns per warp-level load per SM (includes launch + fill overhead)
        cudaFree(out);
        cudaFree(cyc);
        return (double)ms * 1e6 / ((double)iters * 16 * (threads / 32));
    }
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    cudaFree(out);
    cudaFree(cyc);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    double sum = 0;
    for (int i = 0; i < sms; i++) sum += (double)h[i];
    return sum / sms / ((double)iters * 16 * (threads / 32));   // cycles per warp-level load, per SM
}


// Sustained FP32 FMA throughput in TFLOP/s (2 flop per FMA), dependent chains with ILP 8.
UB_API double asw_ubench_ffma_tflops(int packed) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    Arg a{nullptr, 4096, sms * 8, 256, 0};
    cudaMalloc(&a.out, 64);
    float ms;
    if (packed) ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_ffma2<8><<<a->blocks, a->threads>>>(a->out, a->iters, 1.0001f, 0.5f); }, &a, 5);
    else ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_ffma<8><<<a->blocks, a->threads>>>(a->out, a->iters, 1.0001f, 0.5f); }, &a, 5);
    cudaFree(a.out);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    double fmas = (double)a.blocks * a.threads * a.iters * 8 * (packed ? 2 : 1);
    return 2.0 * fmas / (ms * 1e-3) / 1e12;
}

// Shared-memory probe: returns SM-cycles-equivalent: LDS.128 warp-instructions per ns per SM
// for the address pattern `mode` with `fma_per_lds` FFMAs between loads (0, 4, 8, 16).
UB_API double asw_ubench_lds(int mode, int fma_per_lds, double* out_ms) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    Arg a{nullptr, 8192, sms * 4, 512, mode};
    cudaMalloc(&a.out, 64);
    float ms = -1.f;
    switch (fma_per_lds) {
        case 0: ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_lds_mix<0><<<a->blocks, a->threads>>>(a->out, a->iters, a->mode); }, &a, 5); break;
        case 4: ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_lds_mix<4><<<a->blocks, a->threads>>>(a->out, a->iters, a->mode); }, &a, 5); break;
        case 8: ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_lds_mix<8><<<a->blocks, a->threads>>>(a->out, a->iters, a->mode); }, &a, 5); break;
        case 16: ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_lds_mix<16><<<a->blocks, a->threads>>>(a->out, a->iters, a->mode); }, &a, 5); break;
        default: break;
    }
    cudaFree(a.out);
    if (ms < 0 || cudaGetLastError() != cudaSuccess) return -1.0;
    if (out_ms) *out_ms = ms;
    double lds = (double)a.blocks * (a.threads / 32) * a.iters;  // warp-level LDS.128 instructions
    return lds / (ms * 1e6) / sms;                                // per ns per SM
}

// FP32 pipe throughput of tap-shaped instruction mixes, in "tap-ops" (one FMUL + one FMA, scalar
// equivalent) per clock per SM at `clock_mhz`; the roofline is 64 (128 lanes x 1 FMA/clk / 2 per tap).
UB_API double asw_ubench_mix(int mode, double clock_mhz) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static float* in = nullptr;
    if (!in) { cudaMalloc(&in, 256); cudaMemset(in, 0, 256); }
    Arg a{nullptr, 4096, sms * 8, 256, 0};
    cudaMalloc(&a.out, 64);
    static float* gin; gin = in;
    float ms = -1.f;
    switch (mode) {
        case 0: ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_mix<0><<<a->blocks, a->threads>>>(a->out, a->iters, gin); }, &a, 5); break;
        case 1: ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_mix<1><<<a->blocks, a->threads>>>(a->out, a->iters, gin); }, &a, 5); break;
        case 2: ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_mix<2><<<a->blocks, a->threads>>>(a->out, a->iters, gin); }, &a, 5); break;
        case 3: ms = time_ms([](void* p) { Arg* a = (Arg*)p; k_mix<3><<<a->blocks, a->threads>>>(a->out, a->iters, gin); }, &a, 5); break;
        default: break;
    }
    cudaFree(a.out);
    if (ms < 0 || cudaGetLastError() != cudaSuccess) return -1.0;
    // lane-level ops per iteration: mode 0/1: 16 tap-ops; mode 2/3: 16 single instructions (= 8 tap-op equivalents)
    double lane_ops = (double)a.blocks * a.threads * a.iters * 16.0;
    double per_clk_sm = lane_ops / (ms * 1e-3) / (clock_mhz * 1e6) / sms;
    return per_clk_sm;   // lanes-ops per clock per SM: FMA peak = 128 single instr / 64 tap-ops
}
