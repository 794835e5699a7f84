/*
 * asw_oracle.c -- CPU restatement of the reference's ASW hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the reported CPU baseline.
 * The product path (stereo_matchin_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  The restatement is checked in tests/test_oracle_golden.py
 * against the reference's own committed outputs (asw_consistency_pre-reff.png for the
 * hot path, asw_disparity.png for hot path + tail, sukub/asw_raw_d.png for raw cost +
 * WTA) -- see SURVEY.md section 4 / 8(c).  The reference itself cannot run here (it
 * needs an OpenCL device and MSVC CRT calls), so this port is also the CPU baseline
 * ("kind": "port").
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/stereo_matching/).  All arithmetic is float32.  Layouts are the
 * reference's: images RGBA8 tightly packed (main.cpp:189,243); cost volumes
 * x + W*y + W*H*d; support tables x + W*y + W*H*tap.
 *
 * Floating-point freedom of the reference (OpenCL, no build options, main.cpp:211):
 * a*b+c may or may not be contracted to an FMA, exp is <= 3 ulp.  This port fixes
 * both: `use_fma` selects fmaf() or separately rounded mul+add for the accumulation
 * (the file is compiled with -ffp-contract=off so nothing else is contracted), and
 * exp is the correctly rounded float of the double-precision exp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

typedef struct {
    int radius;      /* 16  : kernels/asw_vsupport.cl:19, asw_vcost_aggregation.cl:33,36 */
    int ndisp;       /* 61  : kernels/asw_aggr.cl:16, asw_wta.cl:34 */
    float gamma_c;   /* 30.91f : kernels/asw_vsupport.cl:22 */
    float gamma_p;   /* 28.21f : kernels/asw_vsupport.cl:24 */
    float trunc;     /* +inf: the reference does not truncate (asw_aggr.cl:19) */
    int iterations;  /* 7   : main.cpp:177 (r) */
} oracle_params;

ORACLE_API void oracle_params_default(oracle_params* p) {
    p->radius = 16; p->ndisp = 61; p->gamma_c = 30.91f; p->gamma_p = 28.21f;
    p->trunc = INFINITY; p->iterations = 7;
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* read_imagef(CL_UNORM_INT8) * 255  (asw_aggr.cl:12, asw_vsupport.cl:20): the sampler
 * returns v/255.0f, the kernels multiply back by 255 -> two float roundings. */
static inline float px(uint8_t v) { return (float)v / 255.0f * 255.0f; }

/* RGBA8 -> three float planes (alpha ignored: only .x .y .z are used, asw_aggr.cl:19) */
static void unpack_rgb(const uint8_t* rgba, int W, int H, float* r, float* g, float* b) {
    float lut[256];
    for (int i = 0; i < 256; i++) lut[i] = px((uint8_t)i);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < W * H; i++) {
        r[i] = lut[rgba[4 * i + 0]];
        g[i] = lut[rgba[4 * i + 1]];
        b[i] = lut[rgba[4 * i + 2]];
    }
}

/* write_imagef(float -> CL_UNORM_INT8).  The committed goldens show round-half-down
 * of the float32 product f*255. */
static inline uint8_t q8(float f) {
    float t = f * 255.0f;
    if (!(t > 0.0f)) return 0;
    if (t >= 255.0f) return 255;
    return (uint8_t)ceilf(t - 0.5f);
}

ORACLE_API uint8_t oracle_q8(float f) { return q8(f); }

/* -------------------------------------------------------------------------- */
/* asw_Aggr -- kernels/asw_aggr.cl:3-23.  cost[x,y,d] = SAD_rgb(L(x,y), R(clamp(x-d),y)),
 * 0..255 scale; `trunc` (not in the reference) caps the value, +inf = reference. */
ORACLE_API void oracle_asw_aggr(const uint8_t* left_rgba, const uint8_t* right_rgba,
                                int W, int H, int D, float trunc, float* cost) {
    size_t n = (size_t)W * H;
    float* buf = (float*)malloc(sizeof(float) * n * 6);
    float *lr = buf, *lg = buf + n, *lb = buf + 2 * n, *rr = buf + 3 * n, *rg = buf + 4 * n, *rb = buf + 5 * n;
    unpack_rgb(left_rgba, W, H, lr, lg, lb);
    unpack_rgb(right_rgba, W, H, rr, rg, rb);
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < D; d++) {
        for (int y = 0; y < H; y++) {
            float* out = cost + (size_t)d * n + (size_t)y * W;
            const size_t row = (size_t)y * W;
            for (int x = 0; x < W; x++) {
                int xr = x - d; if (xr < 0) xr = 0;          /* CLAMP_TO_EDGE, main.cpp:10 */
                float v = fabsf(lr[row + x] - rr[row + xr]) + fabsf(lg[row + x] - rg[row + xr])
                        + fabsf(lb[row + x] - rb[row + xr]);   /* asw_aggr.cl:19 */
                out[x] = v < trunc ? v : trunc;
            }
        }
    }
    free(buf);
}

/* exp(): correctly rounded float of the double exp (reference: OpenCL exp, <= 3 ulp) */
static inline float exp_f32(float x) { return (float)exp((double)x); }

/* -------------------------------------------------------------------------- */
/* asw_vSupport / asw_hSupport -- kernels/asw_vsupport.cl:3-27, asw_hsupport.cl:3-28.
 * w[x,y,i] = exp( -SAD(I(x,y), I(q)) / gamma_c - dist(p,q) / gamma_p ), q = p shifted by
 * (i - R) along y (vertical) or x (horizontal) and CLAMPED to the image; the distance
 * uses the clamped coordinate (asw_vsupport.cl:19,24). */
static void support(const uint8_t* rgba, int W, int H, int R, float gc, float gp, int vertical, float* out) {
    size_t n = (size_t)W * H;
    int T = 2 * R + 1;
    float* buf = (float*)malloc(sizeof(float) * n * 3);
    float *cr = buf, *cg = buf + n, *cb = buf + 2 * n;
    unpack_rgb(rgba, W, H, cr, cg, cb);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < T; i++) {
        for (int y = 0; y < H; y++) {
            float* o = out + (size_t)i * n + (size_t)y * W;
            for (int x = 0; x < W; x++) {
                int qx = x, qy = y;
                if (vertical) qy = clampi(y + i - R, 0, H - 1); else qx = clampi(x + i - R, 0, W - 1);
                size_t p = (size_t)y * W + x, q = (size_t)qy * W + qx;
                float sad = fabsf(cr[p] - cr[q]) + fabsf(cg[p] - cg[q]) + fabsf(cb[p] - cb[q]);
                float c_diff = (-sad) / gc;                               /* asw_vsupport.cl:22 */
                float dx = (float)(x - qx), dy = (float)(y - qy);
                float g_dist = sqrtf(dx * dx + dy * dy) / gp;             /* distance(), :24 */
                o[x] = exp_f32(c_diff - g_dist);                          /* :25 */
            }
        }
    }
    free(buf);
}

ORACLE_API void oracle_asw_vsupport(const uint8_t* rgba, int W, int H, int R, float gc, float gp, float* out) {
    support(rgba, W, H, R, gc, gp, 1, out);
}
ORACLE_API void oracle_asw_hsupport(const uint8_t* rgba, int W, int H, int R, float gc, float gp, float* out) {
    support(rgba, W, H, R, gc, gp, 0, out);
}

/* -------------------------------------------------------------------------- */
/* asw_vCostAggregation -- kernels/asw_vcost_aggregation.cl:11-44
 * asw_hCostAggregation -- kernels/asw_hcost_aggregation.cl:12-44
 *   xr = max(x-d,0); num = den = 1e-5f;
 *   for i in 0..T-1 (in order): ww = sL[x,y,i]*sR[xr,y,i]; num += ww*cin[tap]; den += ww;
 *   cout = num/den  (vertical also stores den, :43; horizontal ignores its denom input).
 * Work is arranged row-wise so the inner x loops vectorise; the per-output accumulation
 * order (tap 0..T-1) is exactly the reference's. */
#define AGG_BODY(FMA)                                                                      \
    for (int i = 0; i < T; i++) {                                                          \
        const float* sl = sL + (size_t)i * n + row;                                        \
        const float* sr = sR + (size_t)i * n + row;                                        \
        const float* c;                                                                    \
        int sh = 0;                                                                        \
        if (vertical) c = cin + (size_t)d * n + (size_t)clampi(y + i - R, 0, H - 1) * W;  \
        else { c = cin + (size_t)d * n + row; sh = i - R; }                                \
        int xe = d < W ? d : W;                                                            \
        for (int x = 0; x < xe; x++) { /* x - d < 0 -> column 0 */                         \
            float ww = sl[x] * sr[0];                                                      \
            float cv = vertical ? c[x] : c[clampi(x + sh, 0, W - 1)];                      \
            num[x] = FMA ? fmaf(ww, cv, num[x]) : num[x] + ww * cv;                        \
            den[x] = den[x] + ww;                                                          \
        }                                                                                  \
        if (vertical) {                                                                    \
            for (int x = xe; x < W; x++) {                                                 \
                float ww = sl[x] * sr[x - d];                                              \
                num[x] = FMA ? fmaf(ww, c[x], num[x]) : num[x] + ww * c[x];                \
                den[x] = den[x] + ww;                                                      \
            }                                                                              \
        } else {                                                                           \
            int lo = xe > -sh ? xe : -sh; if (lo > W) lo = W;                              \
            int hi = W - sh < W ? W - sh : W; if (hi < lo) hi = lo;                        \
            for (int x = xe; x < lo; x++) {                                                \
                float ww = sl[x] * sr[x - d];                                              \
                float cv = c[clampi(x + sh, 0, W - 1)];                                    \
                num[x] = FMA ? fmaf(ww, cv, num[x]) : num[x] + ww * cv;                    \
                den[x] = den[x] + ww;                                                      \
            }                                                                              \
            for (int x = lo; x < hi; x++) {                                                \
                float ww = sl[x] * sr[x - d];                                              \
                num[x] = FMA ? fmaf(ww, c[x + sh], num[x]) : num[x] + ww * c[x + sh];      \
                den[x] = den[x] + ww;                                                      \
            }                                                                              \
            for (int x = hi; x < W; x++) {                                                 \
                float ww = sl[x] * sr[x - d];                                              \
                float cv = c[clampi(x + sh, 0, W - 1)];                                    \
                num[x] = FMA ? fmaf(ww, cv, num[x]) : num[x] + ww * cv;                    \
                den[x] = den[x] + ww;                                                      \
            }                                                                              \
        }                                                                                  \
    }

static void aggregate(const float* sL, const float* sR, const float* cin, int W, int H, int D, int R,
                      int vertical, int use_fma, float* denom_out, float* cout) {
    const size_t n = (size_t)W * H;
    const int T = 2 * R + 1;
#pragma omp parallel
    {
        float* num = (float*)malloc(sizeof(float) * W * 2);
        float* den = num + W;
#pragma omp for collapse(2) schedule(static)
        for (int d = 0; d < D; d++) {
            for (int y = 0; y < H; y++) {
                const size_t row = (size_t)y * W;
                for (int x = 0; x < W; x++) { num[x] = 0.00001f; den[x] = 0.00001f; }
                if (use_fma) { AGG_BODY(1) } else { AGG_BODY(0) }
                float* o = cout + (size_t)d * n + row;
                for (int x = 0; x < W; x++) o[x] = num[x] / den[x];
                if (denom_out) memcpy(denom_out + (size_t)d * n + row, den, sizeof(float) * W);
            }
        }
        free(num);
    }
}

ORACLE_API void oracle_asw_vcost_aggregation(const float* supp_left, const float* supp_right, const float* input_cost,
                                             int W, int H, int D, int R, int use_fma,
                                             float* output_denom, float* output_cost) {
    aggregate(supp_left, supp_right, input_cost, W, H, D, R, 1, use_fma, output_denom, output_cost);
}
ORACLE_API void oracle_asw_hcost_aggregation(const float* supp_left, const float* supp_right, const float* vertical_cost,
                                             int W, int H, int D, int R, int use_fma, float* output_cost) {
    aggregate(supp_left, supp_right, vertical_cost, W, H, D, R, 0, use_fma, NULL, output_cost);
}

/* -------------------------------------------------------------------------- */
/* asw_WTA -- kernels/asw_wta.cl:12-82.  Left part (:25-47,70,73,76-77): ascending scan
 * with strict '<' (lowest d wins ties), two-min tracking from sentinels 100000,
 * conf = (min2-min1)/min2.  Right/target part (:50-67,71,74,79-80): for i<min_d sample
 * cost[max(0,x-i), y, b] with b from `bresenham` (:3-9), results stored at (x,y).
 * Any output pointer may be NULL. */
static inline int bresenham(int p1x, int p1y, int p2x, int p2y, int x) {
    int y = p1x;
    if ((p1y - p2y) != 0) y = (p1x - p2x) / (p1y - p2y) * (x - p2y) + p2x;   /* asw_wta.cl:6-7, int division */
    return y;
}

ORACLE_API void oracle_asw_wta(const float* cost, int W, int H, int D,
                               uint8_t* out_rgba, float* d_est_ref, float* d_est_tar, uint8_t* out_tar_rgba,
                               float* conf_ref, float* conf_tar) {
    const size_t n = (size_t)W * H;
    const float scale = (float)(D - 1);                       /* 60.0f, asw_wta.cl:70 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x;
            float cur = 100000.0f, last = 100000.0f;
            int min_d = 0;
            for (int i = 0; i < D; i++) {
                float t = cost[p + n * i];
                last = t < last ? t : last;                   /* :43 */
                min_d = t < cur ? i : min_d;                  /* :44 */
                last = t < cur ? cur : last;                  /* :45 */
                cur = t < cur ? t : cur;                      /* :46 */
            }
            int d_r = min_d, min_d_r = min_d;
            float cur_t = 100000.0f, last_t = 100000.0f;
            if (d_est_tar || out_tar_rgba || conf_tar) {
                for (int i = 0; i < d_r; i++) {
                    int xq = x - i > 0 ? x - i : 0;
                    int b = bresenham(0, x - d_r, min_d, x, xq);          /* :57 */
                    float t = cost[(size_t)xq + (size_t)W * y + n * b];    /* :62 */
                    last_t = t < last_t ? t : last_t;
                    min_d_r = t < cur_t ? b : min_d_r;
                    last_t = t < cur_t ? cur_t : last_t;
                    cur_t = t < cur_t ? t : cur_t;
                }
            }
            if (out_rgba) {
                uint8_t v = D > 1 ? q8((float)min_d / scale) : 0;
                out_rgba[4 * p] = v; out_rgba[4 * p + 1] = v; out_rgba[4 * p + 2] = v; out_rgba[4 * p + 3] = 255;
            }
            if (out_tar_rgba) {
                uint8_t v = D > 1 ? q8((float)min_d_r / scale) : 0;
                out_tar_rgba[4 * p] = v; out_tar_rgba[4 * p + 1] = v; out_tar_rgba[4 * p + 2] = v; out_tar_rgba[4 * p + 3] = 255;
            }
            if (d_est_ref) d_est_ref[p] = (float)min_d;
            if (conf_ref) conf_ref[p] = (last - cur) / last;
            if (d_est_tar) d_est_tar[p] = (float)min_d_r;
            if (conf_tar) conf_tar[p] = (last_t - cur_t) / last_t;
        }
    }
}

/* -------------------------------------------------------------------------- */
/* Constistency -- kernels/consist.cl:3-34 (first consumer of the hot path; used to pin
 * the oracle against asw_consistency_pre-reff.png, main.cpp:531-536,625-627).
 * Images are read back as UNORM8 (v/255) and scaled by `dscale` (60 = D-1). */
ORACLE_API void oracle_consistency(const uint8_t* ref_rgba, const uint8_t* tar_rgba, int W, int H, float dscale,
                                   float* conf_ref, float* conf_tar, uint8_t* out_rgba, uint8_t* out_red_rgba) {
#pragma omp parallel for schedule(static)
    for (int p = 0; p < W * H; p++) {
        int ok0 = 0;
        for (int c = 0; c < 4; c++) {
            float pr = (float)ref_rgba[4 * p + c] / 255.0f * dscale;      /* consist.cl:20 */
            float pt = (float)tar_rgba[4 * p + c] / 255.0f * dscale;      /* :21 */
            int ok = fabsf(pt - pr) < 1.001f;                             /* :25 */
            if (c == 0) ok0 = ok;
            if (out_rgba) out_rgba[4 * p + c] = q8(ok ? pr / dscale : pt / dscale);        /* :30,32 */
            if (out_red_rgba) {
                float red = (c == 0 || c == 3) ? 1.0f : 0.0f;                              /* :23 */
                out_red_rgba[4 * p + c] = q8(ok ? pr / dscale : red);                       /* :25,33 */
            }
        }
        if (conf_ref && !ok0) conf_ref[p] = 0.0f;                         /* :27 */
        if (conf_tar && !ok0) conf_tar[p] = 0.0f;                         /* :28 */
    }
}

/* -------------------------------------------------------------------------- */
/* The host sequence of main.cpp:463-526: raw cost, 4 support tables, r x (V, H), WTA.
 * Optional outputs (NULL to skip): final volume, left/right RGBA maps, d_est, conf.
 * Returns 0, or -1 on allocation failure / bad arguments. */
ORACLE_API int oracle_asw_hot_path(const uint8_t* left_rgba, const uint8_t* right_rgba, int W, int H,
                                   const oracle_params* prm, int use_fma,
                                   float* final_cost, uint8_t* out_rgba, uint8_t* out_tar_rgba,
                                   float* d_est_ref, float* d_est_tar, float* conf_ref, float* conf_tar) {
    if (!left_rgba || !right_rgba || W <= 0 || H <= 0 || !prm || prm->ndisp <= 0 || prm->radius < 0 || prm->iterations < 0)
        return -1;
    const int D = prm->ndisp, R = prm->radius, T = 2 * R + 1;
    const size_t n = (size_t)W * H;
    float* vol0 = (float*)malloc(sizeof(float) * n * D);
    float* vol1 = (float*)malloc(sizeof(float) * n * D);
    float* vol2 = (float*)malloc(sizeof(float) * n * D);
    float* den = (float*)malloc(sizeof(float) * n * D);
    float* vl = (float*)malloc(sizeof(float) * n * T);
    float* hl = (float*)malloc(sizeof(float) * n * T);
    float* vr = (float*)malloc(sizeof(float) * n * T);
    float* hr = (float*)malloc(sizeof(float) * n * T);
    int rc = -1;
    if (vol0 && vol1 && vol2 && den && vl && hl && vr && hr) {
        oracle_asw_aggr(left_rgba, right_rgba, W, H, D, prm->trunc, vol0);              /* main.cpp:463-466 */
        oracle_asw_vsupport(left_rgba, W, H, R, prm->gamma_c, prm->gamma_p, vl);         /* :470-472 */
        oracle_asw_hsupport(left_rgba, W, H, R, prm->gamma_c, prm->gamma_p, hl);         /* :474-476 */
        oracle_asw_vsupport(right_rgba, W, H, R, prm->gamma_c, prm->gamma_p, vr);        /* :478-480 */
        oracle_asw_hsupport(right_rgba, W, H, R, prm->gamma_c, prm->gamma_p, hr);        /* :482-484 */
        const float* in = vol0;
        for (int it = 0; it < prm->iterations; it++) {                                   /* :492-515 */
            oracle_asw_vcost_aggregation(vl, vr, in, W, H, D, R, use_fma, den, vol1);
            oracle_asw_hcost_aggregation(hl, hr, vol1, W, H, D, R, use_fma, vol2);
            in = vol2;
        }
        if (final_cost) memcpy(final_cost, in, sizeof(float) * n * D);
        oracle_asw_wta(in, W, H, D, out_rgba, d_est_ref, d_est_tar, out_tar_rgba, conf_ref, conf_tar);  /* :519-526 */
        rc = 0;
    }
    free(vol0); free(vol1); free(vol2); free(den); free(vl); free(hl); free(vr); free(hr);
    return rc;
}
