// asw_api.cu -- the extern "C" boundary (include/asw_b200.h) and the host sequence of the
// ASW hot path.  Replaces the OpenCL runtime glue and enqueue sequence of the reference's
// stereo_matching/main.cpp:119-130,158-172,210-256,434-526,621 with CUDA: one context =
// one GPU + one stream + grow-on-demand device scratch (allocation stays outside the
// timed region, as in the reference where clCreateBuffer precedes the first event).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/asw_b200.h"
#include "asw_common.cuh"
#include "asw_kernels_basic.cuh"
#include "asw_kernels_tma.cuh"
#include "asw_kernels_tail.cuh"
#include "asw_kernels_cross.cuh"

using namespace asw;

namespace {

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct asw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    LaunchEnv env;                         // SM count + cached tensor maps of this context
    cudaStream_t side = nullptr;            // second stream: boundary rows of the horizontal pass under the interior rows (halo overlap)
    cudaEvent_t ev_v = nullptr, ev_hb = nullptr;
    cudaDeviceProp prop{};
    std::string err;
    int family = 0;
    int keep_volume = 0;
    const float* final_volume = nullptr;
    int launches = 0;
    // device scratch, reused across calls
    Scratch img_l, img_r;                  // uploads (host entry point)
    Scratch out_rgba, out_d, out_conf;     // downloads (host entry point)
    Scratch vL, hL, vR, hR;                // support tables
    Scratch vol[3];                        // cost volumes (raw / ping / pong)
    Scratch den_v, den_h;                  // hoisted denominators
    Scratch wta_part;                      // per-window (min1, min2, argmin) of the winner-take-all fused into the last H pass
    Scratch vol_ref;                       // final volume in the reference layout (keep_volume)
    Scratch fimg_l, fimg_r;                // images as float4 (r, g, b, 0), sampler conversion applied
    Scratch tail[12];                      // whole-method buffers (asw_stereo)
    Scratch cb[10];                        // cross-based method buffers (asw_cross_stereo)
    enum { kMaxEvents = 96 };
    cudaEvent_t ev[kMaxEvents] = {};
    // CUDA graphs of the hot path, one per call signature (buffers, shape, parameters): the second call with a signature is
    // captured, later ones replay the graph (the ~30 launches of a frame become one; it matters for sub-millisecond frames)
    struct GraphKey {
        const void *dl, *dr, *rgba, *dd, *conf;
        int W, H, y0, y1, family, fuse_env;
        asw_params prm;
    };
    enum { kGraphSlots = 4 };
    GraphKey gkey[kGraphSlots] = {};
    cudaGraphExec_t gexec[kGraphSlots] = {};
    int gstate[kGraphSlots] = {};          // 0 empty, 1 signature seen once, 2 graph ready
    int gnext = 0;
    bool graphs_off = false;
};

namespace {

int fail(asw_ctx* c, int status, const char* what, cudaError_t e = cudaSuccess) {
    if (c) {
        char buf[512];
        if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
        else snprintf(buf, sizeof buf, "%s", what);
        c->err = buf;
    }
    return status;
}

#define CU(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fail(ctx, ASW_ERR_CUDA, #call, e__); \
    } while (0)

void drop_graphs(asw_ctx* ctx) {
    for (int i = 0; i < asw_ctx::kGraphSlots; i++) {
        if (ctx->gexec[i]) cudaGraphExecDestroy(ctx->gexec[i]);
        ctx->gexec[i] = nullptr;
        ctx->gstate[i] = 0;
    }
}

int ensure(asw_ctx* ctx, Scratch& s, size_t bytes) {
    if (s.cap >= bytes) return ASW_OK;
    drop_graphs(ctx);                                          // captured launches point into the scratch buffers
    if (s.p) { cudaFree(s.p); s.p = nullptr; s.cap = 0; }
    cudaError_t e = cudaMalloc(&s.p, bytes);
    if (e != cudaSuccess) { s.p = nullptr; cudaGetLastError(); return fail(ctx, ASW_ERR_NOMEM, "cudaMalloc scratch", e); }
    s.cap = bytes;
    // Zero once: padding planes / zero-weight taps may touch elements no kernel has written yet,
    // and 0 * garbage must stay finite.  Later contents are always finite results of earlier runs.
    e = cudaMemsetAsync(s.p, 0, bytes, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, ASW_ERR_CUDA, "cudaMemsetAsync scratch", e);
    return ASW_OK;
}

int check_params(asw_ctx* ctx, int W, int H, const asw_params* p) {
    if (!ctx) return ASW_ERR_INVALID;
    if (!p) return fail(ctx, ASW_ERR_INVALID, "params is NULL");
    if (W <= 0 || H <= 0) return fail(ctx, ASW_ERR_INVALID, "W and H must be positive");
    if (p->ndisp <= 0 || p->radius < 0 || p->iterations < 0) return fail(ctx, ASW_ERR_INVALID, "ndisp > 0, radius >= 0, iterations >= 0 required");
    if (!(p->gamma_c > 0.0f) || !(p->gamma_p > 0.0f)) return fail(ctx, ASW_ERR_INVALID, "gamma_c and gamma_p must be positive");
    // raw costs must be positive and normal for the branch-free division of the fused kernels (div_rn_normal); NaN fails the test too
    if (!(p->trunc > 0.0f)) return fail(ctx, ASW_ERR_INVALID, "trunc must be positive (INFINITY = no truncation)");
    if (p->ndisp > 65535 || H > 65535) return fail(ctx, ASW_ERR_UNSUPPORTED, "ndisp and H are limited to 65535");
    if (p->radius > 64) return fail(ctx, ASW_ERR_UNSUPPORTED, "radius is limited to 64");
    return ASW_OK;
}

inline dim3 grid3(int W, int rows, int z, int bx) { return dim3((unsigned)((W + bx - 1) / bx), (unsigned)rows, (unsigned)z); }

// ---- basic-family launches (also the per-operator entry points) -----------------------
int launch_raw(asw_ctx* ctx, const uint8_t* l, const uint8_t* r, Band b, int ylo, int yhi, const asw_params* p, float* cost) {
    if (yhi <= ylo) return ASW_OK;
    int gz = p->ndisp < 8 ? p->ndisp : 8;
    k_raw_cost<<<grid3(b.W, yhi - ylo, gz, 128), 128, 0, ctx->stream>>>((const uint32_t*)l, (const uint32_t*)r, b, ylo, yhi,
                                                                      p->ndisp, p->trunc, cost);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int launch_support(asw_ctx* ctx, bool vertical, const uint8_t* img, Band b, int ylo, int yhi, const asw_params* p, float* out) {
    if (yhi <= ylo) return ASW_OK;
    dim3 g = grid3(b.W, yhi - ylo, 2 * p->radius + 1, 128);
    if (vertical) k_support<true><<<g, 128, 0, ctx->stream>>>((const uint32_t*)img, b, ylo, yhi, p->radius, p->gamma_c, p->gamma_p, out);
    else k_support<false><<<g, 128, 0, ctx->stream>>>((const uint32_t*)img, b, ylo, yhi, p->radius, p->gamma_c, p->gamma_p, out);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int launch_agg_basic(asw_ctx* ctx, bool vertical, Band b, int ylo, int yhi, const asw_params* p, const float* sL,
                     const float* sR, const float* cin, float* den, float* cout) {
    if (yhi <= ylo) return ASW_OK;
    dim3 g = grid3(b.W, yhi - ylo, p->ndisp, 128);
    if (vertical) k_aggregate<true><<<g, 128, 0, ctx->stream>>>(sL, sR, cin, b, ylo, yhi, p->radius, den, cout);
    else k_aggregate<false><<<g, 128, 0, ctx->stream>>>(sL, sR, cin, b, ylo, yhi, p->radius, den, cout);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int launch_wta(asw_ctx* ctx, Band b, int ylo, int yhi, int out_y0, const asw_params* p, const float* cost, uint8_t* rgba,
               uint8_t* dd, float* d_ref, float* d_tar, uint8_t* tar_rgba, float* conf_ref, float* conf_tar) {
    if (yhi <= ylo) return ASW_OK;
    k_wta<<<grid3(b.W, yhi - ylo, 1, 128), 128, 0, ctx->stream>>>(cost, b, ylo, yhi, p->ndisp, out_y0, (uint32_t*)rgba, dd, d_ref,
                                                                 d_tar, (uint32_t*)tar_rgba, conf_ref, conf_tar);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

struct EvTimer {
    asw_ctx* ctx;
    bool on;
    int n = 0;
    int mark() {
        if (!on || n >= asw_ctx::kMaxEvents) return -1;
        cudaEventRecord(ctx->ev[n], ctx->stream);
        return n++;
    }
    float ms(int a, int b) const {
        if (!on || a < 0 || b < 0) return 0.f;
        float t = 0.f;
        cudaEventElapsedTime(&t, ctx->ev[a], ctx->ev[b]);
        return t;
    }
};

// Records the stage boundaries of one run and turns them into the reference's log columns.
struct StageTimes {
    EvTimer et;
    int e_start = -1, e_raw = -1, e_supp = -1, e_agg = -1, e_wta = -1;
    bool pass_marks = false;
    int ev_v0[24], ev_vmain[24], ev_v1[24], ev_h1[24];   // per iteration: V start, V main kernel end, V end, H end
    int prev = -1;                                       // the latest mark
    void v_begin(int it) { if (pass_marks) { ev_v0[it] = prev; ev_vmain[it] = -1; } }
    void v_end(int it) { if (pass_marks) prev = ev_v1[it] = et.mark(); }
    void h_end(int it) { if (pass_marks) prev = ev_h1[it] = et.mark(); }
    void fill(asw_timing* tm, int r, int launches) {
        memset(tm, 0, sizeof *tm);
        tm->raw_ms = et.ms(e_start, e_raw);
        tm->supp_ms = et.ms(e_raw, e_supp);
        float vsum = 0.f, hsum = 0.f, fsum = 0.f;
        for (int it = 0; pass_marks && it < r; it++) {
            vsum += et.ms(ev_v0[it], ev_v1[it]);
            hsum += et.ms(ev_v1[it], ev_h1[it]);
            if (ev_vmain[it] >= 0) fsum += et.ms(ev_vmain[it], ev_v1[it]);
        }
        tm->vagg_mean_ms = r ? vsum / r : 0.f;
        tm->hagg_mean_ms = r ? hsum / r : 0.f;
        tm->vfix_mean_ms = r ? fsum / r : 0.f;
        tm->agg_total_ms = et.ms(e_supp, e_agg);
        tm->wta_ms = et.ms(e_agg, e_wta);
        tm->total_ms = et.ms(e_start, e_wta);
        tm->kernel_launches = launches;
    }
};

#define CUL(call)                                                                          \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        ctx->launches++;                                                                   \
        if (e__ != cudaSuccess) return fail(ctx, ASW_ERR_CUDA, #call, e__);                \
    } while (0)

// The hot path on device buffers for output rows [y0, y1) of a W x H frame (main.cpp:463-526).
struct Shard {            // disparity shard [d0, d1) of the problem and where its partial WTA result goes (TMA family only)
    int d0 = 0, d1 = -1;
    float *min1 = nullptr, *min2 = nullptr;
    int* arg = nullptr;
};

struct HaloX {            // per-iteration halo exchange with the neighbouring row bands
    asw_halo_fn fn = nullptr;               // host-synchronous exchange (asw_disparity_band_exchange_device)
    asw_halo_begin_fn begin = nullptr;      // or: asynchronous exchange hidden under the interior rows (..._async_device)
    asw_halo_end_fn end = nullptr;
    void* user = nullptr;
};

int run_band_impl(asw_ctx* ctx, const uint8_t* dl, const uint8_t* dr, int W, int H, int y0, int y1, const asw_params* p,
                  uint8_t* d_rgba, uint8_t* d_d, float* d_conf, asw_timing* tm, const Shard* sh, const HaloX* hx) {
    const int R = p->radius, T = 2 * R + 1, Dfull = p->ndisp, r = p->iterations;
    const int sd0 = sh ? sh->d0 : 0, sd1 = sh ? sh->d1 : Dfull;
    const int D = sd1 - sd0;                                   // disparities aggregated by this call
    // rows whose values can influence rows [y0,y1): R rows per V pass (H passes stay in-row).  With a halo exchange the
    // band keeps only R halo rows and every iteration works on exactly [y0,y1): the neighbours' rows arrive between iterations.
    const int ya = max(0, y0 - (hx ? R : r * R)), yb = min(H, y1 + (hx ? R : r * R));
    Band b{W, H, ya, yb - ya};
    const bool tma = ctx->family == 0 && tma_supported(R, Dfull);
    if (sh && !tma) return fail(ctx, ASW_ERR_UNSUPPORTED, "disparity shards need the TMA kernel family (radius 16, family 0)");
    if (hx && !tma) return fail(ctx, ASW_ERR_UNSUPPORTED, "the halo exchange needs the TMA kernel family (radius 16, family 0)");
    const TL tl = make_tl(b, Dfull, sd0, sd1);
    const int Dp = tma ? tma_padded_D(D) : D;
    const size_t vol_bytes = sizeof(float) * (tma ? tl.vol_elems() : b.plane() * (size_t)Dp);
    const size_t tab_bytes = sizeof(float) * b.plane() * (size_t)T;
    int st;
    if (tma) {
        if ((st = ensure(ctx, ctx->fimg_l, sizeof(float) * 4 * (size_t)W * H)) || (st = ensure(ctx, ctx->fimg_r, sizeof(float) * 4 * (size_t)W * H))) return st;
        if ((st = ensure(ctx, ctx->vL, sizeof(float) * tl.wvl_elems())) || (st = ensure(ctx, ctx->vR, sizeof(float) * tl.wvr_elems())) ||
            (st = ensure(ctx, ctx->hL, sizeof(float) * tl.whl_elems())) || (st = ensure(ctx, ctx->hR, sizeof(float) * tl.whr_elems())))
            return st;
    } else if ((st = ensure(ctx, ctx->vL, tab_bytes)) || (st = ensure(ctx, ctx->hL, tab_bytes)) ||
               (st = ensure(ctx, ctx->vR, tab_bytes)) || (st = ensure(ctx, ctx->hR, tab_bytes)))
        return st;
    for (int i = 0; i < (tma ? 2 : 3); i++)
        if ((st = ensure(ctx, ctx->vol[i], vol_bytes))) return st;
    // the TMA family keeps the vertical denominators in a private per-thread layout (vden_* in asw_kernels_tma.cuh)
    const size_t denv_bytes = tma ? sizeof(float) * vden_total_floats(W, b.y_off, b.Hb, Dp) : vol_bytes;
    if (tma && r > 0 && ((st = ensure(ctx, ctx->den_v, denv_bytes)) || (st = ensure(ctx, ctx->den_h, vol_bytes)))) return st;
    float *vL = (float*)ctx->vL.p, *hL = (float*)ctx->hL.p, *vR = (float*)ctx->vR.p, *hR = (float*)ctx->hR.p;

    ctx->launches = 0;
    ctx->final_volume = nullptr;
    StageTimes t{EvTimer{ctx, tm != nullptr}};
    t.pass_marks = r <= 24;
    t.e_start = t.et.mark();
    const float* fin = nullptr;
    if (tma) {
        float *va = (float*)ctx->vol[0].p, *vb = (float*)ctx->vol[1].p;
        float *den_v = (float*)ctx->den_v.p, *den_h = (float*)ctx->den_h.p;
        cudaStream_t s = ctx->stream;
        const float4 *fl = (const float4*)ctx->fimg_l.p, *fr = (const float4*)ctx->fimg_r.p;
        CUL(launch_unpack_v2(s, dl, W * H, (float4*)ctx->fimg_l.p));
        CUL(launch_unpack_v2(s, dr, W * H, (float4*)ctx->fimg_r.p));
        CUL(launch_raw_v2(s, fl, fr, tl, ya, yb, p->trunc, va));
        t.e_raw = t.et.mark();
        const int ws0 = hx ? y0 : ya, ws1 = hx ? y1 : yb;     // rows that are ever aggregated: only they need weights
        CUL(launch_support4_v2(s, fl, fr, tl, ws0, ws1, p->gamma_c, p->gamma_p, vL, hL, vR, hR));   // main.cpp:470-484, one launch
        t.prev = t.e_supp = t.et.mark();
        // Halo exchange hidden under the interior rows (hx->begin / hx->end): an iteration computes the boundary rows of
        // its horizontal pass first (side stream, concurrently with the interior rows on the main stream), hands them to the
        // transfer, and the next iteration aggregates the interior rows of its vertical pass (which read no halo row) before it
        // waits for the neighbours' rows and finishes the R-row borders.  Split points of the vertical pass are multiples of
        // 8 rows (its tiles), so no tile is shared between two launches.
        const bool up = y0 > 0, down = y1 < H;                    // neighbours
        const int vA = up ? min(y1, (y0 + R + 7) & ~7) : y0, vB = down ? max(vA, (y1 - R) & ~7) : y1;
        const bool overlap = hx && hx->begin && (up || down) && vB - vA >= 8 && y1 - y0 >= 2 * R;
        const size_t vrow = (size_t)tl.Wv * tl.Dp, hbytes = sizeof(float) * vrow * R;
        // Winner-take-all inside the epilogue of the last horizontal pass (asw_wta.cl:25-47 on the values still in registers):
        // the final volume is then never written.  Not when the caller keeps the volume (right view, refinement) or asks for a
        // disparity shard's partial result.
        const int nwin = tl.Dp / (tl.Dp % 128 == 0 ? 128 : 64);
        const bool fuse_wta = r > 0 && hagg_can_fuse_wta() && !ctx->keep_volume && !sh && !(getenv("ASW_FUSE_WTA") && atoi(getenv("ASW_FUSE_WTA")) == 0);
        const size_t npx = (size_t)W * (y1 - y0);
        HWtaOut wo{nullptr, nullptr, nullptr, y0, npx};
        if (fuse_wta) {
            if ((st = ensure(ctx, ctx->wta_part, sizeof(float) * 3 * (size_t)nwin * npx))) return st;
            wo.min1 = (float*)ctx->wta_part.p;
            wo.min2 = wo.min1 + (size_t)nwin * npx;
            wo.arg = (int*)(wo.min2 + (size_t)nwin * npx);
        }
        for (int it = 0; it < r; it++) {
            const int ylo = hx ? y0 : max(ya, y0 - (r - 1 - it) * R), yhi = hx ? y1 : min(yb, y1 + (r - 1 - it) * R);
            cudaEvent_t ev_main = nullptr;
            t.v_begin(it);
            if (t.pass_marks && t.et.on && t.et.n < asw_ctx::kMaxEvents) {   // an event between the main kernel and its fix-up / padding launches
                t.ev_vmain[it] = t.et.n++;
                ev_main = ctx->ev[t.ev_vmain[it]];
            }
            if (overlap && it > 0) {
                CUL(launch_vagg_v2(s, false, tl, vA, vB, vL, vR, va, den_v, vb, ev_main, &ctx->env));
                if (hx->end(hx->user, it - 1, (void*)s)) return fail(ctx, ASW_ERR_CUDA, "halo exchange (end) callback failed");
                if (vA > y0) { CUL(launch_vagg_v2(s, false, tl, y0, vA, vL, vR, va, den_v, vb, nullptr, &ctx->env)); ctx->launches += 2; }
                if (y1 > vB) { CUL(launch_vagg_v2(s, false, tl, vB, y1, vL, vR, va, den_v, vb, nullptr, &ctx->env)); ctx->launches += 2; }
            } else {
                CUL(launch_vagg_v2(s, it == 0, tl, ylo, yhi, vL, vR, va, den_v, vb, ev_main, &ctx->env));
            }
            ctx->launches += kVHelpers ? 1 : 2;                // main kernel (+ diagonal fix-up kernel) + edge padding kernel
            t.v_end(it);
            if (overlap && it + 1 < r) {
                const int h0 = up ? y0 + R : y0, h1 = down ? y1 - R : y1;       // interior rows of the horizontal pass
                CU(cudaEventRecord(ctx->ev_v, s));
                CU(cudaStreamWaitEvent(ctx->side, ctx->ev_v, 0));
                if (up) CUL(launch_hagg_v2(ctx->side, it == 0, tl, y0, h0, hL, hR, vb, den_h, va, &ctx->env));
                if (down) CUL(launch_hagg_v2(ctx->side, it == 0, tl, h1, y1, hL, hR, vb, den_h, va, &ctx->env));
                CU(cudaEventRecord(ctx->ev_hb, ctx->side));
                if (hx->begin(hx->user, it, up ? va + (size_t)(y0 - ya) * vrow : nullptr, down ? va + (size_t)(y1 - R - ya) * vrow : nullptr,
                              up ? va + (size_t)(y0 - R - ya) * vrow : nullptr, down ? va + (size_t)(y1 - ya) * vrow : nullptr, hbytes,
                              (void*)ctx->side))
                    return fail(ctx, ASW_ERR_CUDA, "halo exchange (begin) callback failed");
                CUL(launch_hagg_v2(s, it == 0, tl, h0, h1, hL, hR, vb, den_h, va, &ctx->env));
                CU(cudaStreamWaitEvent(s, ctx->ev_hb, 0));     // the next vertical pass reads the boundary rows as well
                t.h_end(it);
                continue;
            }
            CUL(launch_hagg_v2(s, it == 0, tl, ylo, yhi, hL, hR, vb, den_h, va, &ctx->env, fuse_wta && it + 1 == r ? &wo : nullptr));
            t.h_end(it);
            if (hx && hx->fn && it + 1 < r) {
                // the next vertical pass reads R rows of each neighbour: hand out our boundary rows (volume rows are contiguous:
                // Wv * Dp floats each) and where the neighbours' rows go; the callback moves them (NCCL, peer copies ...)
                const int hs = hx->fn(hx->user, it, up ? va + (size_t)(y0 - ya) * vrow : nullptr, down ? va + (size_t)(y1 - R - ya) * vrow : nullptr,
                                      up ? va + (size_t)(y0 - R - ya) * vrow : nullptr, down ? va + (size_t)(y1 - ya) * vrow : nullptr, hbytes);
                if (hs) return fail(ctx, ASW_ERR_CUDA, "halo exchange callback failed");
                t.prev = t.et.mark();                          // the exchange counts towards agg_total_ms, not towards the next V pass
            } else if (hx && hx->begin && !overlap && it + 1 < r) {
                // band too short to split: the asynchronous callbacks are used back to back (no overlap)
                if (hx->begin(hx->user, it, up ? va + (size_t)(y0 - ya) * vrow : nullptr, down ? va + (size_t)(y1 - R - ya) * vrow : nullptr,
                              up ? va + (size_t)(y0 - R - ya) * vrow : nullptr, down ? va + (size_t)(y1 - ya) * vrow : nullptr, hbytes, (void*)s) ||
                    hx->end(hx->user, it, (void*)s))
                    return fail(ctx, ASW_ERR_CUDA, "halo exchange callback failed");
            }
        }
        t.e_agg = t.et.mark();
        if (fuse_wta) {
            k_wta_merge<<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(wo.min1, wo.min2, wo.arg, nwin, npx, Dfull, (uint32_t*)d_rgba, d_d, d_conf);
            CUL(cudaGetLastError());
        } else {
            CUL(launch_wta_v2(s, tl, y0, y1, y0, Dfull, va, d_rgba, d_d, d_conf, sh ? sh->min1 : nullptr, sh ? sh->min2 : nullptr, sh ? sh->arg : nullptr));
        }
        t.e_wta = t.et.mark();
        if (ctx->keep_volume) {
            if ((st = ensure(ctx, ctx->vol_ref, sizeof(float) * (size_t)W * (y1 - y0) * D))) return st;
            CUL(launch_volume_to_ref_v2(s, tl, y0, y1, va, (float*)ctx->vol_ref.p));
            fin = (const float*)ctx->vol_ref.p;
        }
    } else {
        float *raw = (float*)ctx->vol[0].p, *va = (float*)ctx->vol[1].p, *hb = (float*)ctx->vol[2].p;
        if ((st = launch_raw(ctx, dl, dr, b, ya, yb, p, raw))) return st;                 // main.cpp:463-466
        t.e_raw = t.et.mark();
        if ((st = launch_support(ctx, true, dl, b, ya, yb, p, vL))) return st;            // main.cpp:470-472
        if ((st = launch_support(ctx, false, dl, b, ya, yb, p, hL))) return st;           // :474-476
        if ((st = launch_support(ctx, true, dr, b, ya, yb, p, vR))) return st;            // :478-480
        if ((st = launch_support(ctx, false, dr, b, ya, yb, p, hR))) return st;           // :482-484
        t.prev = t.e_supp = t.et.mark();
        const float* in = raw;
        for (int it = 0; it < r; it++) {                                                  // main.cpp:492-515
            const int ylo = max(ya, y0 - (r - 1 - it) * R), yhi = min(yb, y1 + (r - 1 - it) * R);
            t.v_begin(it);
            if ((st = launch_agg_basic(ctx, true, b, ylo, yhi, p, vL, vR, in, nullptr, va))) return st;
            t.v_end(it);
            if ((st = launch_agg_basic(ctx, false, b, ylo, yhi, p, hL, hR, va, nullptr, hb))) return st;
            t.h_end(it);
            in = hb;
        }
        t.e_agg = t.et.mark();
        if ((st = launch_wta(ctx, b, y0, y1, y0, p, in, d_rgba, d_d, nullptr, nullptr, nullptr, d_conf, nullptr))) return st;  // main.cpp:519-526
        t.e_wta = t.et.mark();
        if (ctx->keep_volume) {
            // band-local rows [ya,yb) -> hand out rows [y0,y1) only when the band is the whole buffer
            if (ya == y0 && yb == y1) fin = in;
            else {
                if ((st = ensure(ctx, ctx->vol_ref, sizeof(float) * (size_t)W * (y1 - y0) * D))) return st;
                // rows [y0,y1) of every disparity plane: one strided copy (source pitch = a band plane, destination pitch = the rows kept)
                CU(cudaMemcpy2DAsync(ctx->vol_ref.p, sizeof(float) * (size_t)W * (y1 - y0), in + (size_t)(y0 - ya) * W, sizeof(float) * b.plane(),
                                     sizeof(float) * (size_t)W * (y1 - y0), (size_t)D, cudaMemcpyDeviceToDevice, ctx->stream));
                fin = (const float*)ctx->vol_ref.p;
            }
        }
    }
    ctx->final_volume = fin;
    if (tm) {
        CU(cudaStreamSynchronize(ctx->stream));
        t.fill(tm, r, ctx->launches);
    }
    return ASW_OK;
}

// run_band_impl, or the CUDA graph captured from it.  A plain call (TMA family, no timing, no kept volume, no shard, no halo
// exchange) whose signature -- buffers, shape, band, parameters -- repeats is captured on its second occurrence and replayed
// from then on: one graph launch instead of ~30 kernel launches (frames of the reference's size take ~1 ms).
// Any reallocation of a scratch buffer drops the graphs (ensure()); ASW_GRAPH=0 switches the mechanism off.
int run_band(asw_ctx* ctx, const uint8_t* dl, const uint8_t* dr, int W, int H, int y0, int y1, const asw_params* p,
             uint8_t* d_rgba, uint8_t* d_d, float* d_conf, asw_timing* tm, const Shard* sh = nullptr, const HaloX* hx = nullptr) {
    static const bool env_off = getenv("ASW_GRAPH") && atoi(getenv("ASW_GRAPH")) == 0;
    const bool eligible = !env_off && !ctx->graphs_off && !tm && !sh && !hx && !ctx->keep_volume && ctx->family == 0 &&
                          tma_supported(p->radius, p->ndisp) && p->iterations > 0;
    if (!eligible) return run_band_impl(ctx, dl, dr, W, H, y0, y1, p, d_rgba, d_d, d_conf, tm, sh, hx);
    asw_ctx::GraphKey k;
    memset(&k, 0, sizeof k);
    k.dl = dl; k.dr = dr; k.rgba = d_rgba; k.dd = d_d; k.conf = d_conf;
    k.W = W; k.H = H; k.y0 = y0; k.y1 = y1; k.family = ctx->family;
    k.fuse_env = getenv("ASW_FUSE_WTA") ? atoi(getenv("ASW_FUSE_WTA")) : -1;
    k.prm = *p;
    int slot = -1;
    for (int i = 0; i < asw_ctx::kGraphSlots; i++)
        if (ctx->gstate[i] && !memcmp(&ctx->gkey[i], &k, sizeof k)) slot = i;
    if (slot >= 0 && ctx->gstate[slot] == 2) {
        cudaError_t e = cudaGraphLaunch(ctx->gexec[slot], ctx->stream);
        if (e != cudaSuccess) return fail(ctx, ASW_ERR_CUDA, "cudaGraphLaunch", e);
        ctx->final_volume = nullptr;
        return ASW_OK;
    }
    if (slot < 0) {                                            // first occurrence: remember the signature, run normally
        slot = ctx->gnext;
        ctx->gnext = (ctx->gnext + 1) % asw_ctx::kGraphSlots;
        if (ctx->gexec[slot]) { cudaGraphExecDestroy(ctx->gexec[slot]); ctx->gexec[slot] = nullptr; }
        memcpy(&ctx->gkey[slot], &k, sizeof k);
        ctx->gstate[slot] = 1;
        return run_band_impl(ctx, dl, dr, W, H, y0, y1, p, d_rgba, d_d, d_conf, tm, sh, hx);
    }
    // second occurrence: every scratch buffer already has its size, so the launch sequence can be captured
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        ctx->graphs_off = true;
        return run_band_impl(ctx, dl, dr, W, H, y0, y1, p, d_rgba, d_d, d_conf, tm, sh, hx);
    }
    const int st = run_band_impl(ctx, dl, dr, W, H, y0, y1, p, d_rgba, d_d, d_conf, tm, sh, hx);
    const cudaError_t ec = cudaStreamEndCapture(ctx->stream, &graph);
    // the slot may have been cleared by ensure() during the capture (a buffer grew after all): then nothing was cached
    if (st != ASW_OK || ec != cudaSuccess || !graph || ctx->gstate[slot] != 1 || memcmp(&ctx->gkey[slot], &k, sizeof k)) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        ctx->graphs_off = true;                                // be conservative: this context keeps launching kernel by kernel
        ctx->gstate[slot] = 0;
        return run_band_impl(ctx, dl, dr, W, H, y0, y1, p, d_rgba, d_d, d_conf, tm, sh, hx);
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess || !exec) {
        cudaGetLastError();
        ctx->graphs_off = true;
        ctx->gstate[slot] = 0;
        return run_band_impl(ctx, dl, dr, W, H, y0, y1, p, d_rgba, d_d, d_conf, tm, sh, hx);
    }
    ctx->gexec[slot] = exec;
    ctx->gstate[slot] = 2;
    const cudaError_t el = cudaGraphLaunch(exec, ctx->stream);
    if (el != cudaSuccess) return fail(ctx, ASW_ERR_CUDA, "cudaGraphLaunch", el);
    ctx->final_volume = nullptr;
    return ASW_OK;
}

}  // namespace

extern "C" {

const char* asw_version(void) { return "asw_b200 0.1 (sm_100a)"; }

const char* asw_strerror(int s) {
    switch (s) {
        case ASW_OK: return "ok";
        case ASW_ERR_INVALID: return "invalid argument";
        case ASW_ERR_CUDA: return "CUDA runtime error";
        case ASW_ERR_NOMEM: return "out of memory";
        case ASW_ERR_UNSUPPORTED: return "unsupported parameter";
        default: return "unknown status";
    }
}

void asw_params_default(asw_params* p) {
    if (!p) return;
    p->radius = 16; p->ndisp = 61; p->gamma_c = 30.91f; p->gamma_p = 28.21f; p->trunc = INFINITY; p->iterations = 7;
}

int asw_create(asw_ctx** out, int device) {
    if (!out) return ASW_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return ASW_ERR_CUDA; }   // no GPU: fail loudly, no fallback
    if (device < 0 || device >= n) return ASW_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return ASW_ERR_CUDA;
    asw_ctx* ctx = new (std::nothrow) asw_ctx();
    if (!ctx) return ASW_ERR_NOMEM;
    ctx->device = device;
    if (cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_v, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_hb, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return ASW_ERR_CUDA;
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    ctx->env.sms = ctx->prop.multiProcessorCount;
    if (tma_configure() != cudaSuccess) { cudaGetLastError(); asw_destroy(ctx); return ASW_ERR_CUDA; }
    *out = ctx;
    return ASW_OK;
}

int asw_destroy(asw_ctx* ctx) {
    if (!ctx) return ASW_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drop_graphs(ctx);
    Scratch* all[] = {&ctx->img_l, &ctx->img_r, &ctx->out_rgba, &ctx->out_d, &ctx->out_conf, &ctx->vL, &ctx->hL, &ctx->vR,
                      &ctx->hR, &ctx->vol[0], &ctx->vol[1], &ctx->vol[2], &ctx->den_v, &ctx->den_h, &ctx->wta_part, &ctx->vol_ref, &ctx->fimg_l, &ctx->fimg_r};
    for (Scratch* s : all) if (s->p) cudaFree(s->p);
    for (Scratch& s : ctx->tail) if (s.p) cudaFree(s.p);
    for (Scratch& s : ctx->cb) if (s.p) cudaFree(s.p);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    cudaStreamSynchronize(ctx->side);
    if (ctx->ev_v) cudaEventDestroy(ctx->ev_v);
    if (ctx->ev_hb) cudaEventDestroy(ctx->ev_hb);
    cudaStreamDestroy(ctx->side);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ASW_OK;
}

const char* asw_last_error(asw_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* asw_stream(asw_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int asw_sync(asw_ctx* ctx) {
    if (!ctx) return ASW_ERR_INVALID;
    CU(cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

int asw_device_info(asw_ctx* ctx, int* sm_count, int* sm_clock_khz, size_t* total_mem, char* name, size_t name_len) {
    if (!ctx) return ASW_ERR_INVALID;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (sm_clock_khz) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device); *sm_clock_khz = khz; }
    if (total_mem) *total_mem = ctx->prop.totalGlobalMem;
    if (name && name_len) { strncpy(name, ctx->prop.name, name_len - 1); name[name_len - 1] = 0; }
    return ASW_OK;
}

int asw_set_kernel_family(asw_ctx* ctx, int family) {
    if (!ctx || family < 0 || family > 1) return ASW_ERR_INVALID;
    ctx->family = family;
    return ASW_OK;
}

int asw_set_keep_volume(asw_ctx* ctx, int keep) {
    if (!ctx) return ASW_ERR_INVALID;
    ctx->keep_volume = keep != 0;
    return ASW_OK;
}

const float* asw_final_volume(asw_ctx* ctx) { return ctx ? ctx->final_volume : nullptr; }

int asw_disparity_band_device(asw_ctx* ctx, const uint8_t* dl, const uint8_t* dr, int W, int H, int y0, int y1,
                              const asw_params* prm, uint8_t* d_rgba, uint8_t* d_d, float* d_conf, asw_timing* tm) {
    int st = check_params(ctx, W, H, prm);
    if (st) return st;
    if (!dl || !dr) return fail(ctx, ASW_ERR_INVALID, "image pointer is NULL");
    if (y0 < 0 || y1 > H || y0 >= y1) return fail(ctx, ASW_ERR_INVALID, "band must satisfy 0 <= y0 < y1 <= H");
    if (prm->ndisp > 256 && d_d) return fail(ctx, ASW_ERR_UNSUPPORTED, "disp_d is uint8: ndisp <= 256 required");
    CU(cudaSetDevice(ctx->device));
    return run_band(ctx, dl, dr, W, H, y0, y1, prm, d_rgba, d_d, d_conf, tm);
}

int asw_disparity_band_exchange_device(asw_ctx* ctx, const uint8_t* dl, const uint8_t* dr, int W, int H, int y0, int y1,
                                       const asw_params* prm, uint8_t* d_rgba, uint8_t* d_d, float* d_conf, asw_halo_fn exchange,
                                       void* user, asw_timing* tm) {
    int st = check_params(ctx, W, H, prm);
    if (st) return st;
    if (!dl || !dr) return fail(ctx, ASW_ERR_INVALID, "image pointer is NULL");
    if (!exchange) return fail(ctx, ASW_ERR_INVALID, "exchange callback is NULL");
    if (y0 < 0 || y1 > H || y0 >= y1) return fail(ctx, ASW_ERR_INVALID, "band must satisfy 0 <= y0 < y1 <= H");
    if ((y0 > 0 || y1 < H) && y1 - y0 < prm->radius) return fail(ctx, ASW_ERR_INVALID, "a band with neighbours needs at least `radius` rows");
    if (prm->ndisp > 256 && d_d) return fail(ctx, ASW_ERR_UNSUPPORTED, "disp_d is uint8: ndisp <= 256 required");
    CU(cudaSetDevice(ctx->device));
    HaloX hx;
    hx.fn = exchange; hx.user = user;
    const int keep = ctx->keep_volume;
    ctx->keep_volume = 0;
    st = run_band(ctx, dl, dr, W, H, y0, y1, prm, d_rgba, d_d, d_conf, tm, nullptr, &hx);
    ctx->keep_volume = keep;
    return st;
}

int asw_disparity_band_exchange_async_device(asw_ctx* ctx, const uint8_t* dl, const uint8_t* dr, int W, int H, int y0, int y1,
                                             const asw_params* prm, uint8_t* d_rgba, uint8_t* d_d, float* d_conf,
                                             asw_halo_begin_fn begin, asw_halo_end_fn end, void* user, asw_timing* tm) {
    int st = check_params(ctx, W, H, prm);
    if (st) return st;
    if (!dl || !dr) return fail(ctx, ASW_ERR_INVALID, "image pointer is NULL");
    if (!begin || !end) return fail(ctx, ASW_ERR_INVALID, "exchange callback is NULL");
    if (y0 < 0 || y1 > H || y0 >= y1) return fail(ctx, ASW_ERR_INVALID, "band must satisfy 0 <= y0 < y1 <= H");
    if ((y0 > 0 || y1 < H) && y1 - y0 < prm->radius) return fail(ctx, ASW_ERR_INVALID, "a band with neighbours needs at least `radius` rows");
    if (prm->ndisp > 256 && d_d) return fail(ctx, ASW_ERR_UNSUPPORTED, "disp_d is uint8: ndisp <= 256 required");
    CU(cudaSetDevice(ctx->device));
    HaloX hx;
    hx.begin = begin; hx.end = end; hx.user = user;
    const int keep = ctx->keep_volume;
    ctx->keep_volume = 0;
    st = run_band(ctx, dl, dr, W, H, y0, y1, prm, d_rgba, d_d, d_conf, tm, nullptr, &hx);
    ctx->keep_volume = keep;
    return st;
}

void* asw_side_stream(asw_ctx* ctx) { return ctx ? (void*)ctx->side : nullptr; }

int asw_disparity_shard_device(asw_ctx* ctx, const uint8_t* dl, const uint8_t* dr, int W, int H, int y0, int y1, int d0, int d1,
                               const asw_params* prm, float* d_min1, float* d_min2, int* d_arg, asw_timing* tm) {
    int st = check_params(ctx, W, H, prm);
    if (st) return st;
    if (!dl || !dr || !d_min1 || !d_min2 || !d_arg) return fail(ctx, ASW_ERR_INVALID, "device pointer is NULL");
    if (y0 < 0 || y1 > H || y0 >= y1) return fail(ctx, ASW_ERR_INVALID, "band must satisfy 0 <= y0 < y1 <= H");
    if (d0 < 0 || d1 > prm->ndisp || d0 >= d1 || d0 % 64 != 0) return fail(ctx, ASW_ERR_INVALID, "shard must satisfy 0 <= d0 < d1 <= ndisp, d0 a multiple of 64");
    CU(cudaSetDevice(ctx->device));
    Shard sh;
    sh.d0 = d0; sh.d1 = d1; sh.min1 = d_min1; sh.min2 = d_min2; sh.arg = d_arg;
    const int keep = ctx->keep_volume;
    ctx->keep_volume = 0;                                      // a shard's volume is partial: nothing to hand out
    st = run_band(ctx, dl, dr, W, H, y0, y1, prm, nullptr, nullptr, nullptr, tm, &sh);
    ctx->keep_volume = keep;
    return st;
}

int asw_merge_shards(asw_ctx* ctx, int W, int rows, int ndisp, int nshards, const float* d_min1, const float* d_min2, const int* d_arg,
                     uint8_t* d_rgba, uint8_t* d_d, float* d_conf) {
    if (!ctx) return ASW_ERR_INVALID;
    if (W <= 0 || rows <= 0 || ndisp <= 0 || nshards <= 0) return fail(ctx, ASW_ERR_INVALID, "W, rows, ndisp, nshards must be positive");
    if (!d_min1 || !d_min2 || !d_arg) return fail(ctx, ASW_ERR_INVALID, "device pointer is NULL");
    if (ndisp > 256 && d_d) return fail(ctx, ASW_ERR_UNSUPPORTED, "disp_d is uint8: ndisp <= 256 required");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)W * rows;
    k_wta_merge<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_min1, d_min2, d_arg, nshards, n, ndisp, (uint32_t*)d_rgba, d_d, d_conf);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int asw_disparity_device(asw_ctx* ctx, const uint8_t* dl, const uint8_t* dr, int W, int H, const asw_params* prm,
                         uint8_t* d_rgba, uint8_t* d_d, float* d_conf, asw_timing* tm) {
    return asw_disparity_band_device(ctx, dl, dr, W, H, 0, H, prm, d_rgba, d_d, d_conf, tm);
}

static int disparity_host(asw_ctx* ctx, const uint8_t* left, const uint8_t* right, int W, int H, const asw_params* prm,
                          uint8_t* disp_rgba, uint8_t* disp_d, float* conf, asw_timing* tm, bool sync);

int asw_disparity(asw_ctx* ctx, const uint8_t* left, const uint8_t* right, int W, int H, const asw_params* prm,
                  uint8_t* disp_rgba, uint8_t* disp_d, float* conf, asw_timing* tm) {
    return disparity_host(ctx, left, right, W, H, prm, disp_rgba, disp_d, conf, tm, true);
}

int asw_disparity_async(asw_ctx* ctx, const uint8_t* left, const uint8_t* right, int W, int H, const asw_params* prm,
                        uint8_t* disp_rgba, uint8_t* disp_d, float* conf) {
    return disparity_host(ctx, left, right, W, H, prm, disp_rgba, disp_d, conf, nullptr, false);
}

static int disparity_host(asw_ctx* ctx, const uint8_t* left, const uint8_t* right, int W, int H, const asw_params* prm,
                          uint8_t* disp_rgba, uint8_t* disp_d, float* conf, asw_timing* tm, bool sync) {
    int st = check_params(ctx, W, H, prm);
    if (st) return st;
    if (!left || !right) return fail(ctx, ASW_ERR_INVALID, "image pointer is NULL");
    CU(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)W * H;
    if ((st = ensure(ctx, ctx->img_l, npx * 4)) || (st = ensure(ctx, ctx->img_r, npx * 4))) return st;
    if (disp_rgba && (st = ensure(ctx, ctx->out_rgba, npx * 4))) return st;
    if (disp_d && (st = ensure(ctx, ctx->out_d, npx))) return st;
    if (conf && (st = ensure(ctx, ctx->out_conf, npx * 4))) return st;
    cudaEvent_t e0 = ctx->ev[asw_ctx::kMaxEvents - 1], e1 = ctx->ev[asw_ctx::kMaxEvents - 2],
                e2 = ctx->ev[asw_ctx::kMaxEvents - 3], e3 = ctx->ev[asw_ctx::kMaxEvents - 4];
    if (tm) cudaEventRecord(e0, ctx->stream);
    CU(cudaMemcpyAsync(ctx->img_l.p, left, npx * 4, cudaMemcpyHostToDevice, ctx->stream));     // main.cpp:243
    CU(cudaMemcpyAsync(ctx->img_r.p, right, npx * 4, cudaMemcpyHostToDevice, ctx->stream));    // main.cpp:244
    if (tm) cudaEventRecord(e1, ctx->stream);
    asw_timing inner;
    st = asw_disparity_band_device(ctx, (const uint8_t*)ctx->img_l.p, (const uint8_t*)ctx->img_r.p, W, H, 0, H, prm,
                                   disp_rgba ? (uint8_t*)ctx->out_rgba.p : nullptr, disp_d ? (uint8_t*)ctx->out_d.p : nullptr,
                                   conf ? (float*)ctx->out_conf.p : nullptr, tm ? &inner : nullptr);
    if (st) return st;
    if (tm) cudaEventRecord(e2, ctx->stream);
    if (disp_rgba) CU(cudaMemcpyAsync(disp_rgba, ctx->out_rgba.p, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));  // main.cpp:621
    if (disp_d) CU(cudaMemcpyAsync(disp_d, ctx->out_d.p, npx, cudaMemcpyDeviceToHost, ctx->stream));
    if (conf) CU(cudaMemcpyAsync(conf, ctx->out_conf.p, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (tm) cudaEventRecord(e3, ctx->stream);
    if (!sync) return ASW_OK;                                  // asw_disparity_async: the caller waits with asw_sync
    CU(cudaStreamSynchronize(ctx->stream));
    if (tm) {
        *tm = inner;
        cudaEventElapsedTime(&tm->h2d_ms, e0, e1);
        cudaEventElapsedTime(&tm->d2h_ms, e2, e3);
    }
    return ASW_OK;
}

// ---- per-operator entry points -----------------------------------------------------------
#define OP_PROLOGUE(ptr_ok)                                               \
    int st = check_params(ctx, W, H, prm);                                \
    if (st) return st;                                                    \
    if (!(ptr_ok)) return fail(ctx, ASW_ERR_INVALID, "device pointer is NULL"); \
    CU(cudaSetDevice(ctx->device));                                       \
    Band b{W, H, 0, H};

}  // extern "C"

namespace {
// For the reference's window (radius 16) and the default kernel family the operators run on the TMA-fed sm_100a kernels:
// inputs in the reference layouts are re-laid out into the kernels' own layouts on the way in (k_pack_support,
// k_ref_to_volume_v2), results on the way out (k_volume_to_ref_v2, k_vden_to_ref).  Any other radius, or
// asw_set_kernel_family(ctx, 1), uses the generic one-thread-per-output kernels (asw_kernels_basic.cuh).
bool op_tma(asw_ctx* ctx, const asw_params* prm) { return ctx->family == 0 && tma_supported(prm->radius, prm->ndisp); }

int op_unpack(asw_ctx* ctx, const uint8_t* img, int W, int H, Scratch& dst) {
    int st = ensure(ctx, dst, sizeof(float) * 4 * (size_t)W * H);
    if (st) return st;
    CUL(launch_unpack_v2(ctx->stream, img, W * H, (float4*)dst.p));
    return ASW_OK;
}

int op_ref_to_volume(asw_ctx* ctx, const TL& tl, const float* ref, float* vol) {
    dim3 blk(32, 32), grd((tl.W + 31) / 32, (tl.Dp + 31) / 32, tl.H);
    k_ref_to_volume_v2<<<grd, blk, 0, ctx->stream>>>(ref, tl, vol);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

template <bool VERTICAL, bool RIGHT>
int op_pack_support(asw_ctx* ctx, const TL& tl, const float* ref, float* out) {
    const int ncols = VERTICAL ? (RIGHT ? tl.WR4 : tl.WL4) : (RIGHT ? tl.NCB * 32 : tl.NXB * 32);
    k_pack_support<VERTICAL, RIGHT><<<dim3((ncols + 127) / 128, tl.H), 128, 0, ctx->stream>>>(ref, tl, out);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

}  // namespace

extern "C" {

int asw_Aggr(asw_ctx* ctx, const uint8_t* l, const uint8_t* r, int W, int H, const asw_params* prm, float* cost) {
    OP_PROLOGUE(l && r && cost)
    if (!op_tma(ctx, prm)) return launch_raw(ctx, l, r, b, 0, H, prm, cost);
    const TL tl = make_tl(b, prm->ndisp);
    if ((st = op_unpack(ctx, l, W, H, ctx->fimg_l)) || (st = op_unpack(ctx, r, W, H, ctx->fimg_r)) ||
        (st = ensure(ctx, ctx->vol[0], sizeof(float) * tl.vol_elems())))
        return st;
    CUL(launch_raw_v2(ctx->stream, (const float4*)ctx->fimg_l.p, (const float4*)ctx->fimg_r.p, tl, 0, H, prm->trunc, (float*)ctx->vol[0].p));
    CUL(launch_volume_to_ref_v2(ctx->stream, tl, 0, H, (const float*)ctx->vol[0].p, cost));
    return ASW_OK;
}

static int op_support(asw_ctx* ctx, bool vertical, const uint8_t* img, int W, int H, const asw_params* prm, float* out) {
    int st = op_unpack(ctx, img, W, H, ctx->fimg_l);
    if (st) return st;
    dim3 grd((W + 127) / 128, H);
    if (vertical) k_support_ref<true><<<grd, 128, 0, ctx->stream>>>((const float4*)ctx->fimg_l.p, W, H, prm->gamma_c, prm->gamma_p, out);
    else k_support_ref<false><<<grd, 128, 0, ctx->stream>>>((const float4*)ctx->fimg_l.p, W, H, prm->gamma_c, prm->gamma_p, out);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int asw_vSupport(asw_ctx* ctx, const uint8_t* img, int W, int H, const asw_params* prm, float* out) {
    OP_PROLOGUE(img && out)
    if (!op_tma(ctx, prm)) return launch_support(ctx, true, img, b, 0, H, prm, out);
    return op_support(ctx, true, img, W, H, prm, out);
}

int asw_hSupport(asw_ctx* ctx, const uint8_t* img, int W, int H, const asw_params* prm, float* out) {
    OP_PROLOGUE(img && out)
    if (!op_tma(ctx, prm)) return launch_support(ctx, false, img, b, 0, H, prm, out);
    return op_support(ctx, false, img, W, H, prm, out);
}

int asw_vCostAggregation(asw_ctx* ctx, int W, int H, const asw_params* prm, const float* sl, const float* sr,
                         const float* cin, float* den, float* cout) {
    OP_PROLOGUE(sl && sr && cin && cout)
    if (!op_tma(ctx, prm)) return launch_agg_basic(ctx, true, b, 0, H, prm, sl, sr, cin, den, cout);
    const TL tl = make_tl(b, prm->ndisp);
    if ((st = ensure(ctx, ctx->vL, sizeof(float) * tl.wvl_elems())) || (st = ensure(ctx, ctx->vR, sizeof(float) * tl.wvr_elems())) ||
        (st = ensure(ctx, ctx->vol[0], sizeof(float) * tl.vol_elems())) || (st = ensure(ctx, ctx->vol[1], sizeof(float) * tl.vol_elems())) ||
        (st = ensure(ctx, ctx->den_v, sizeof(float) * vden_total_floats(W, 0, H, tl.Dp))))
        return st;
    float *vL = (float*)ctx->vL.p, *vR = (float*)ctx->vR.p, *va = (float*)ctx->vol[0].p, *vb = (float*)ctx->vol[1].p, *dv = (float*)ctx->den_v.p;
    if ((st = op_pack_support<true, false>(ctx, tl, sl, vL)) || (st = op_pack_support<true, true>(ctx, tl, sr, vR)) ||
        (st = op_ref_to_volume(ctx, tl, cin, va)))
        return st;
    CUL(launch_vagg_v2(ctx->stream, true, tl, 0, H, vL, vR, va, dv, vb, nullptr, &ctx->env));   // asw_vcost_aggregation.cl:11-44
    ctx->launches += kVHelpers ? 1 : 2;
    CUL(launch_volume_to_ref_v2(ctx->stream, tl, 0, H, vb, cout));
    if (den) {                                                                                  // output_denom (:43)
        k_vden_to_ref<<<dim3((W + 127) / 128, H, prm->ndisp), 128, 0, ctx->stream>>>(dv, tl, den);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return ASW_OK;
}

int asw_hCostAggregation(asw_ctx* ctx, int W, int H, const asw_params* prm, const float* sl, const float* sr,
                         const float* vcost, const float* denom_v, float* cout) {
    (void)denom_v;  // accepted and ignored, as in asw_hcost_aggregation.cl:17
    OP_PROLOGUE(sl && sr && vcost && cout)
    if (!op_tma(ctx, prm)) return launch_agg_basic(ctx, false, b, 0, H, prm, sl, sr, vcost, nullptr, cout);
    const TL tl = make_tl(b, prm->ndisp);
    if ((st = ensure(ctx, ctx->hL, sizeof(float) * tl.whl_elems())) || (st = ensure(ctx, ctx->hR, sizeof(float) * tl.whr_elems())) ||
        (st = ensure(ctx, ctx->vol[0], sizeof(float) * tl.vol_elems())) || (st = ensure(ctx, ctx->vol[1], sizeof(float) * tl.vol_elems())) ||
        (st = ensure(ctx, ctx->den_h, sizeof(float) * tl.vol_elems())))
        return st;
    float *hL = (float*)ctx->hL.p, *hR = (float*)ctx->hR.p, *va = (float*)ctx->vol[0].p, *vb = (float*)ctx->vol[1].p;
    if ((st = op_pack_support<false, false>(ctx, tl, sl, hL)) || (st = op_pack_support<false, true>(ctx, tl, sr, hR)) ||
        (st = op_ref_to_volume(ctx, tl, vcost, va)))
        return st;
    k_vpad_v2<<<H, 256, 0, ctx->stream>>>(tl, va, 0, H);                                        // CLAMP_TO_EDGE columns of the input
    ctx->launches++;
    CUL(launch_hagg_v2(ctx->stream, true, tl, 0, H, hL, hR, va, (float*)ctx->den_h.p, vb, &ctx->env));   // asw_hcost_aggregation.cl:12-44
    CUL(launch_volume_to_ref_v2(ctx->stream, tl, 0, H, vb, cout));
    return ASW_OK;
}

int asw_WTA(asw_ctx* ctx, int W, int H, const asw_params* prm, const float* cost, uint8_t* out_rgba, float* d_ref,
            float* d_tar, uint8_t* out_tar_rgba, float* conf_ref, float* conf_tar) {
    OP_PROLOGUE(cost)
    if (!op_tma(ctx, prm)) return launch_wta(ctx, b, 0, H, 0, prm, cost, out_rgba, nullptr, d_ref, d_tar, out_tar_rgba, conf_ref, conf_tar);
    if (out_rgba || d_ref || conf_ref) {                       // left part (asw_wta.cl:25-47): warp-shuffle two-minimum scan
        const TL tl = make_tl(b, prm->ndisp);
        if ((st = ensure(ctx, ctx->vol[0], sizeof(float) * tl.vol_elems())) || (st = op_ref_to_volume(ctx, tl, cost, (float*)ctx->vol[0].p))) return st;
        CUL(launch_wta_v2(ctx->stream, tl, 0, H, 0, prm->ndisp, (const float*)ctx->vol[0].p, out_rgba, nullptr, conf_ref, nullptr, nullptr, nullptr, d_ref));
    }
    if (d_tar || out_tar_rgba || conf_tar)                     // right / target part (:50-67) walks a data-dependent diagonal of the reference layout
        return launch_wta(ctx, b, 0, H, 0, prm, cost, nullptr, nullptr, nullptr, d_tar, out_tar_rgba, nullptr, conf_tar);
    return ASW_OK;
}

// ---- consumers of the hot path -----------------------------------------------------------------
int asw_Constistency(asw_ctx* ctx, int W, int H, const asw_params* prm, const uint8_t* ref, const uint8_t* tar, float* conf_ref,
                     float* conf_tar, uint8_t* out, uint8_t* out_red) {
    OP_PROLOGUE(ref && tar)
    (void)b;
    const int n = W * H;
    k_consistency<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const uint32_t*)ref, (const uint32_t*)tar, n, (float)(prm->ndisp - 1), conf_ref,
                                                          conf_tar, (uint32_t*)out, (uint32_t*)out_red);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int asw_ref_v(asw_ctx* ctx, int W, int H, const asw_params* prm, const uint8_t* img, const uint8_t* est, const float* conf, float* out) {
    OP_PROLOGUE(img && est && conf && out)
    (void)b;
    k_ref_v<<<(unsigned)((W + 127) / 128) * H, 128, 0, ctx->stream>>>((const uint32_t*)img, (const uint32_t*)est, conf, W, H, prm->radius,
                                                             (float)(prm->ndisp - 1), out);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int asw_ref_h(asw_ctx* ctx, int W, int H, const asw_params* prm, const uint8_t* img, const float* conf, const float* in, float* out) {
    OP_PROLOGUE(img && conf && in && out)
    (void)b;
    k_ref_h<<<(unsigned)((W + 127) / 128) * H, 128, 0, ctx->stream>>>((const uint32_t*)img, conf, in, W, H, prm->radius, out);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int asw_WTA_REF(asw_ctx* ctx, int W, int H, const asw_params* prm, const float* agg, const float* ref, const float* ref_t, uint8_t* out,
                uint8_t* out_t, float* disp_ref, float* disp_ref_t, float* confidence, float* confidence_target) {
    (void)confidence_target;   // never written, as in asw_wta_ref.cl:63,66
    OP_PROLOGUE(agg && ref && ref_t && out && out_t && confidence)
    (void)b;
    k_wta_ref<<<(unsigned)((W + 127) / 128) * H, 128, 0, ctx->stream>>>(agg, ref, ref_t, W, H, prm->ndisp, (uint32_t*)out, (uint32_t*)out_t, disp_ref,
                                                               disp_ref_t, confidence);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

int asw_Median(asw_ctx* ctx, int W, int H, const uint8_t* in, uint8_t* out) {
    if (!ctx) return ASW_ERR_INVALID;
    if (W <= 0 || H <= 0) return fail(ctx, ASW_ERR_INVALID, "bad image size");
    if (!in || !out) return fail(ctx, ASW_ERR_INVALID, "device pointer is NULL");
    CU(cudaSetDevice(ctx->device));
    k_median<<<(unsigned)((W + 127) / 128) * H, 128, 0, ctx->stream>>>((const uint32_t*)in, W, H, (uint32_t*)out);
    ctx->launches++;
    CU(cudaGetLastError());
    return ASW_OK;
}

// The whole method, main.cpp:463-631: fused hot path -> right-view WTA -> consistency -> k x (ref_v L,R;
// ref_h L,R; WTA_REF; consistency) -> median.
int asw_stereo(asw_ctx* ctx, const uint8_t* left, const uint8_t* right, int W, int H, const asw_params* prm, int refine_iters,
               uint8_t* disparity, uint8_t* pre, uint8_t* post, asw_timing* tm, asw_tail_timing* tail) {
    int st = check_params(ctx, W, H, prm);
    if (st) return st;
    if (!left || !right) return fail(ctx, ASW_ERR_INVALID, "image pointer is NULL");
    if (refine_iters < 0) return fail(ctx, ASW_ERR_INVALID, "refine_iters must be >= 0");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)W * H;
    if ((st = ensure(ctx, ctx->img_l, n * 4)) || (st = ensure(ctx, ctx->img_r, n * 4))) return st;
    // 0 left_wta 1 right_wta 2 conf_ref 3 conf_tar 4 consistency 5 red 6 red_reff 7 vref_l 8 vref_r 9 href_l 10 href_r 11 final
    const size_t sz[12] = {n * 4, n * 4, n * 4, n * 4, n * 4, n * 4, n * 4, n * 8, n * 8, n * 8, n * 8, n * 4};
    for (int i = 0; i < 12; i++)
        if ((st = ensure(ctx, ctx->tail[i], sz[i]))) return st;
    uint8_t *lw = (uint8_t*)ctx->tail[0].p, *rw = (uint8_t*)ctx->tail[1].p, *ce = (uint8_t*)ctx->tail[4].p, *red = (uint8_t*)ctx->tail[5].p,
            *red2 = (uint8_t*)ctx->tail[6].p, *fin = (uint8_t*)ctx->tail[11].p;
    float *cr = (float*)ctx->tail[2].p, *ct = (float*)ctx->tail[3].p, *vl = (float*)ctx->tail[7].p, *vr = (float*)ctx->tail[8].p,
          *hl = (float*)ctx->tail[9].p, *hr = (float*)ctx->tail[10].p;
    const uint8_t *dl = (const uint8_t*)ctx->img_l.p, *dr = (const uint8_t*)ctx->img_r.p;
    CU(cudaMemcpyAsync(ctx->img_l.p, left, n * 4, cudaMemcpyHostToDevice, ctx->stream));      // main.cpp:243
    CU(cudaMemcpyAsync(ctx->img_r.p, right, n * 4, cudaMemcpyHostToDevice, ctx->stream));     // main.cpp:244
    const int keep = ctx->keep_volume;
    ctx->keep_volume = 1;                                          // WTA_REF and the right view read the final volume
    asw_timing hot;
    st = run_band(ctx, dl, dr, W, H, 0, H, prm, nullptr, nullptr, nullptr, (tm || tail) ? &hot : nullptr);
    ctx->keep_volume = keep;
    if (st) return st;
    if (tm) *tm = hot;
    const float* vol = ctx->final_volume;
    // the hot path's events were read by run_band (it synchronised): the tail reuses them
    EvTimer et{ctx, tail != nullptr && 3 + 6 * refine_iters + 1 <= asw_ctx::kMaxEvents};
    const int e0 = et.mark();
    // left + right view and both confidences (asw_wta.cl, main.cpp:519-526)
    if ((st = asw_WTA(ctx, W, H, prm, vol, lw, nullptr, nullptr, rw, cr, ct))) return st;
    const int e_wta = et.mark();
    if ((st = asw_Constistency(ctx, W, H, prm, lw, rw, cr, ct, ce, red))) return st;          // main.cpp:531-536
    const int e_cons = et.mark();
    for (int i = 0; i < refine_iters; i++) {                                                  // main.cpp:545-614
        if ((st = asw_ref_v(ctx, W, H, prm, dl, ce, cr, vl))) return st;
        et.mark();
        if ((st = asw_ref_v(ctx, W, H, prm, dr, rw, ct, vr))) return st;
        et.mark();
        if ((st = asw_ref_h(ctx, W, H, prm, dl, cr, vl, hl))) return st;
        et.mark();
        if ((st = asw_ref_h(ctx, W, H, prm, dr, ct, vr, hr))) return st;
        et.mark();
        if ((st = asw_WTA_REF(ctx, W, H, prm, vol, hl, hr, lw, rw, nullptr, nullptr, cr, ct))) return st;
        et.mark();
        if ((st = asw_Constistency(ctx, W, H, prm, lw, rw, cr, ct, ce, red2))) return st;
        et.mark();
    }
    const int e_ref = e_cons + 6 * refine_iters;
    if ((st = asw_Median(ctx, W, H, ce, fin))) return st;                                     // main.cpp:617-619
    const int e_med = et.mark();
    if (disparity) CU(cudaMemcpyAsync(disparity, fin, n * 4, cudaMemcpyDeviceToHost, ctx->stream));                 // main.cpp:621
    if (pre) CU(cudaMemcpyAsync(pre, red, n * 4, cudaMemcpyDeviceToHost, ctx->stream));                             // main.cpp:625
    if (post) CU(cudaMemcpyAsync(post, refine_iters ? red2 : red, n * 4, cudaMemcpyDeviceToHost, ctx->stream));     // main.cpp:629
    CU(cudaStreamSynchronize(ctx->stream));
    if (tm) tm->kernel_launches = ctx->launches;
    if (tail) {
        memset(tail, 0, sizeof *tail);
        if (et.on) {
            tail->right_wta_ms = et.ms(e0, e_wta);
            tail->consistency_ms = et.ms(e_wta, e_cons);
            float sum[6] = {0, 0, 0, 0, 0, 0};
            for (int i = 0; i < refine_iters; i++)
                for (int k = 0; k < 6; k++) sum[k] += et.ms(e_cons + 6 * i + k, e_cons + 6 * i + k + 1);
            const float inv = refine_iters ? 1.0f / refine_iters : 0.f;
            tail->vref_mean_l_ms = sum[0] * inv;
            tail->vref_mean_r_ms = sum[1] * inv;
            tail->href_mean_l_ms = sum[2] * inv;
            tail->href_mean_r_ms = sum[3] * inv;
            tail->wta_ref_mean_ms = sum[4] * inv;
            tail->consistency_mean_ms = sum[5] * inv;
            tail->refinement_total_ms = et.ms(e_cons, e_ref);
            tail->median_ms = et.ms(e_ref, e_med);
            tail->total_ms = hot.total_ms + et.ms(e0, e_med);
        }
    }
    return ASW_OK;
}

// ---- the cross-based method (main.cpp:258-367) ---------------------------------------------------
void asw_cross_params_default(asw_cross_params* p) {
    if (!p) return;
    p->ndisp = 61;
    p->max_arm = 25;
    p->median_local = 3;
}

namespace {
int check_cross(asw_ctx* ctx, int W, int H, const asw_cross_params* p) {
    if (!ctx) return ASW_ERR_INVALID;
    if (!p) return fail(ctx, ASW_ERR_INVALID, "params is NULL");
    if (W <= 0 || H <= 0) return fail(ctx, ASW_ERR_INVALID, "W and H must be positive");
    if (p->ndisp <= 0 || p->max_arm < 1 || p->median_local < 1) return fail(ctx, ASW_ERR_INVALID, "ndisp > 0, max_arm >= 1, median_local >= 1 required");
    if (p->ndisp > 256 || H > 65535) return fail(ctx, ASW_ERR_UNSUPPORTED, "ndisp is limited to 256 (8-bit disparity images), H to 65535");
    return ASW_OK;
}
#define CB_PROLOGUE(ptr_ok)                                                        \
    int st = check_cross(ctx, W, H, prm);                                          \
    if (st) return st;                                                             \
    if (!(ptr_ok)) return fail(ctx, ASW_ERR_INVALID, "device pointer is NULL");    \
    CU(cudaSetDevice(ctx->device));
#define CB_EPILOGUE     \
    ctx->launches++;    \
    CU(cudaGetLastError()); \
    return ASW_OK;
}  // namespace

int asw_Median_grid(asw_ctx* ctx, int W, int H, int local, const uint8_t* in, uint8_t* out) {
    if (ctx && local < 1) return fail(ctx, ASW_ERR_INVALID, "local must be >= 1");
    int st = asw_Median(ctx, W, H, in, out);
    if (st) return st;
    const int We = local * (W / local), He = local * (H / local);
    if (We < W || He < H) {
        k_cb_zero_border<<<dim3((W + 127) / 128, H), 128, 0, ctx->stream>>>((uint32_t*)out, W, H, We, He);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return ASW_OK;
}

int asw_Cross(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const uint8_t* in, int* out) {
    CB_PROLOGUE(in && out)
    k_cb_cross<<<dim3((W + 127) / 128, H), 128, 0, ctx->stream>>>((const uint32_t*)in, W, H, prm->max_arm, out);
    CB_EPILOGUE
}

int asw_Aggregation(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const uint8_t* l, const uint8_t* r, float* cost) {
    CB_PROLOGUE(l && r && cost)
    k_cb_aggregation<<<dim3((W + 127) / 128, H, prm->ndisp), 128, 0, ctx->stream>>>((const uint32_t*)l, (const uint32_t*)r, W, H, cost);
    CB_EPILOGUE
}

int asw_Integral_h(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, float* cost) {
    CB_PROLOGUE(cost)
    const int nrows = H * prm->ndisp;
    k_cb_integral_h<<<(nrows + 31) / 32, dim3(32, 8), 0, ctx->stream>>>(cost, W, nrows);
    CB_EPILOGUE
}

int asw_Integral_v(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, float* cost) {
    CB_PROLOGUE(cost)
    k_cb_integral_v<<<dim3((W + 63) / 64, prm->ndisp), 64, 0, ctx->stream>>>(cost, W, H);
    CB_EPILOGUE
}

int asw_Oii_hcross(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const int* cl, const int* cr, const float* cost, float* tmp) {
    CB_PROLOGUE(cl && cr && cost && tmp)
    k_cb_oii<true><<<dim3((W + 127) / 128, H, prm->ndisp), 128, 0, ctx->stream>>>(cl, cr, cost, W, H, tmp);
    CB_EPILOGUE
}

int asw_Oii_vcross(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const int* cl, const int* cr, const float* tmp, float* cost) {
    CB_PROLOGUE(cl && cr && cost && tmp)
    k_cb_oii<false><<<dim3((W + 127) / 128, H, prm->ndisp), 128, 0, ctx->stream>>>(cl, cr, tmp, W, H, cost);
    CB_EPILOGUE
}

int asw_Init_disparity(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const float* cost, uint8_t* out) {
    CB_PROLOGUE(cost && out)
    k_cb_init_disparity<<<dim3((W + 127) / 128, H), 128, 0, ctx->stream>>>(cost, W, H, prm->ndisp, (uint32_t*)out);
    CB_EPILOGUE
}

int asw_Disparity(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const uint8_t* in, const int* cross, uint8_t* out) {
    CB_PROLOGUE(in && cross && out)
    k_cb_disparity<256><<<dim3((W + 7) / 8, H), dim3(32, 8), 0, ctx->stream>>>((const uint32_t*)in, cross, W, H, prm->ndisp, (uint32_t*)out);
    CB_EPILOGUE
}

int asw_cross_stereo(asw_ctx* ctx, const uint8_t* left, const uint8_t* right, int W, int H, const asw_cross_params* prm, uint8_t* initial,
                     uint8_t* disparity, uint8_t* median_l, asw_cross_timing* tm) {
    int st = check_cross(ctx, W, H, prm);
    if (st) return st;
    if (!left || !right) return fail(ctx, ASW_ERR_INVALID, "image pointer is NULL");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)W * H, vol = sizeof(float) * n * prm->ndisp;
    // 0 left 1 right 2 median_l 3 median_r 4 cross_l 5 cross_r 6 cost 7 temp_cost 8 disparity + f_disparity 9 cross_method
    const size_t sz[10] = {n * 4, n * 4, n * 4, n * 4, n * 16, n * 16, vol, vol, n * 8, n * 4};
    for (int i = 0; i < 10; i++)
        if ((st = ensure(ctx, ctx->cb[i], sz[i]))) return st;
    uint8_t *dl = (uint8_t*)ctx->cb[0].p, *dr = (uint8_t*)ctx->cb[1].p, *ml = (uint8_t*)ctx->cb[2].p, *mr = (uint8_t*)ctx->cb[3].p;
    int *cl = (int*)ctx->cb[4].p, *cr = (int*)ctx->cb[5].p;
    float *cost = (float*)ctx->cb[6].p, *tmp = (float*)ctx->cb[7].p;
    uint8_t *disp = (uint8_t*)ctx->cb[8].p, *fdisp = disp + n * 4, *fin = (uint8_t*)ctx->cb[9].p;
    ctx->launches = 0;
    CU(cudaMemcpyAsync(dl, left, n * 4, cudaMemcpyHostToDevice, ctx->stream));                 // main.cpp:243-244
    CU(cudaMemcpyAsync(dr, right, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    EvTimer et{ctx, tm != nullptr};
    const int e0 = et.mark();
    if ((st = asw_Median_grid(ctx, W, H, prm->median_local, dl, ml))) return st;               // :270-274
    const int e1 = et.mark();
    if ((st = asw_Median_grid(ctx, W, H, prm->median_local, dr, mr))) return st;               // :276-279
    const int e2 = et.mark();
    if ((st = asw_Cross(ctx, W, H, prm, ml, cl))) return st;                                   // :283-286
    const int e3 = et.mark();
    if ((st = asw_Cross(ctx, W, H, prm, mr, cr))) return st;                                   // :288-291
    const int e4 = et.mark();
    if ((st = asw_Aggregation(ctx, W, H, prm, ml, mr, cost))) return st;                       // :295-299
    const int e5 = et.mark();
    if ((st = asw_Integral_h(ctx, W, H, prm, cost))) return st;                                // :304-307
    const int e6 = et.mark();
    if ((st = asw_Oii_hcross(ctx, W, H, prm, cl, cr, cost, tmp))) return st;                   // :312-318
    const int e7 = et.mark();
    if ((st = asw_Integral_v(ctx, W, H, prm, tmp))) return st;                                 // :322-325
    const int e8 = et.mark();
    if ((st = asw_Oii_vcross(ctx, W, H, prm, cl, cr, tmp, cost))) return st;                   // :329-335
    const int e9 = et.mark();
    if ((st = asw_Init_disparity(ctx, W, H, prm, cost, disp))) return st;                      // :339-342
    const int e10 = et.mark();
    if ((st = asw_Disparity(ctx, W, H, prm, disp, cl, fdisp))) return st;                      // :346-350
    const int e11 = et.mark();
    if ((st = asw_Median_grid(ctx, W, H, prm->median_local, fdisp, fin))) return st;           // :352-354
    const int e12 = et.mark();
    if (initial) CU(cudaMemcpyAsync(initial, disp, n * 4, cudaMemcpyDeviceToHost, ctx->stream));     // :357-359
    if (disparity) CU(cudaMemcpyAsync(disparity, fin, n * 4, cudaMemcpyDeviceToHost, ctx->stream));  // :361-363
    if (median_l) CU(cudaMemcpyAsync(median_l, ml, n * 4, cudaMemcpyDeviceToHost, ctx->stream));     // :365-367
    CU(cudaStreamSynchronize(ctx->stream));
    if (tm) {
        memset(tm, 0, sizeof *tm);
        tm->median_l_ms = et.ms(e0, e1);  tm->median_r_ms = et.ms(e1, e2);  tm->median_ms = et.ms(e0, e2);
        tm->cross_l_ms = et.ms(e2, e3);   tm->cross_r_ms = et.ms(e3, e4);   tm->cross_ms = et.ms(e2, e4);
        tm->aggregation_ms = et.ms(e4, e5);
        tm->integral_h_ms = et.ms(e5, e6);  tm->oii_h_ms = et.ms(e6, e7);
        tm->integral_v_ms = et.ms(e7, e8);  tm->oii_v_ms = et.ms(e8, e9);
        tm->init_disparity_ms = et.ms(e9, e10);
        tm->final_disparity_ms = et.ms(e10, e11);
        tm->total_ms = et.ms(e0, e12);
    }
    return ASW_OK;
}

// ---- memory helpers --------------------------------------------------------------------------
int asw_dev_alloc(asw_ctx* ctx, void** p, size_t bytes) {
    if (!ctx || !p) return ASW_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess) { *p = nullptr; cudaGetLastError(); return fail(ctx, ASW_ERR_NOMEM, "cudaMalloc", e); }
    return ASW_OK;
}

int asw_dev_free(asw_ctx* ctx, void* p) {
    if (!ctx) return ASW_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(p));
    return ASW_OK;
}

int asw_memcpy_h2d(asw_ctx* ctx, void* d, const void* h, size_t bytes) {
    if (!ctx || (!d && bytes) || (!h && bytes)) return ASW_ERR_INVALID;
    CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

int asw_memcpy_d2h(asw_ctx* ctx, void* h, const void* d, size_t bytes) {
    if (!ctx || (!d && bytes) || (!h && bytes)) return ASW_ERR_INVALID;
    CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

int asw_host_alloc(asw_ctx* ctx, void** p, size_t bytes) {
    if (!ctx || !p) return ASW_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { *p = nullptr; cudaGetLastError(); return fail(ctx, ASW_ERR_NOMEM, "cudaHostAlloc", e); }
    return ASW_OK;
}

int asw_host_free(asw_ctx* ctx, void* p) {
    if (!ctx) return ASW_ERR_INVALID;
    CU(cudaFreeHost(p));
    return ASW_OK;
}

#ifdef ASW_VPROF
// instrumented builds only: the vertical pass' private denominator buffer of the last call (scripts/race_probe.py)
__attribute__((visibility("default"))) long long asw_debug_vden(asw_ctx* ctx, float* out_host, long long max_floats) {
    cudaDeviceSynchronize();
    long long n = (long long)(ctx->den_v.cap / sizeof(float));
    if (n > max_floats) n = max_floats;
    if (out_host && n > 0) cudaMemcpy(out_host, ctx->den_v.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
    return n;
}
// instrumented builds only (scripts/vprof.py): reads and clears the vertical pass' phase counters
__attribute__((visibility("default"))) int asw_debug_vprof(unsigned long long* out16) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, asw::g_vprof, sizeof(unsigned long long) * 16);
    static const unsigned long long zero[16] = {0};
    cudaMemcpyToSymbol(asw::g_vprof, zero, sizeof zero);
    return 0;
}
#endif
}  // extern "C"
