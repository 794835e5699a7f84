/*
 * asw_tail_oracle.c -- CPU restatement of the consumers of the ASW hot path ("next" rows of
 * SURVEY.md section 8f): iterative disparity refinement, penalised WTA and the final median,
 * plus the whole-method driver of main.cpp:529-631.
 *
 * TEST INFRASTRUCTURE ONLY (see the header of asw_oracle.c).  Parity status: PINNED against the
 * reference's committed <dataset>/asw_disparity.png (tests/test_oracle_golden.py); every quirk of
 * the reference that those PNGs depend on is kept and marked QUIRK.
 * Paths are relative to /root/reference/stereo_matching/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

typedef struct {
    int radius, ndisp;
    float gamma_c, gamma_p, trunc;
    int iterations;
} oracle_params;

/* from asw_oracle.c */
uint8_t oracle_q8(float f);
int oracle_asw_hot_path(const uint8_t*, const uint8_t*, int, int, const oracle_params*, int, float*, uint8_t*, uint8_t*, float*, float*,
                        float*, float*);
void oracle_consistency(const uint8_t*, const uint8_t*, int, int, float, float*, float*, uint8_t*, uint8_t*);

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float px(uint8_t v) { return (float)v / 255.0f * 255.0f; }                 /* read_imagef * 255 */
static inline float exp_f32(float x) { return (float)exp((double)x); }

/* supp_v / supp_h -- kernels/asw_refinement_v.cl:2-10, asw_refinement_h.cl:2-12:
 * exp(-SAD/10.94 - dist/118.78) with the clamped-coordinate distance */
static inline float supp(const uint8_t* p, const uint8_t* q, int dist) {
    float sad = fabsf(px(p[0]) - px(q[0])) + fabsf(px(p[1]) - px(q[1])) + fabsf(px(p[2]) - px(q[2]));
    float c_diff = (-sad) / 10.94f;
    float g_dist = (float)dist / 118.78f;
    return exp_f32(c_diff - g_dist);
}

/* asw_ref_v -- kernels/asw_refinement_v.cl:13-51.  est is an RGBA8 disparity image read as
 * v/255 * dscale (dscale = 60 in the reference, :38); out has two planes: num/den and den. */
ORACLE_API void oracle_asw_ref_v(const uint8_t* img, const uint8_t* est, const float* conf, int W, int H, int R, float dscale,
                                 int use_fma, float* out) {
    const size_t n = (size_t)W * H;
#pragma omp parallel for collapse(2) schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const uint8_t* p = img + 4 * ((size_t)y * W + x);
            float num = 0.00001f, den = 0.00001f;
            for (int i = 0; i < 2 * R + 1; i++) {
                int yy = clampi(y + i - R, 0, H - 1);
                size_t qi = (size_t)yy * W + x;
                float Dx = (float)est[4 * qi] / 255.0f * dscale;            /* :38 */
                float ww = supp(p, img + 4 * qi, abs(y - yy));               /* :39 */
                float F = conf[qi];                                           /* :40 */
                float wf = ww * F;
                if (use_fma) { num = fmaf(wf, Dx, num); den = fmaf(ww, F, den); }
                else { num = num + wf * Dx; den = den + wf; }                 /* :42-43 */
            }
            out[(size_t)y * W + x] = num / den;                               /* :49 */
            out[(size_t)y * W + x + n] = den;                                 /* :50 */
        }
}

/* asw_ref_h -- kernels/asw_refinement_h.cl:16-53.  in = vertical result (2 planes). */
ORACLE_API void oracle_asw_ref_h(const uint8_t* img, const float* conf, const float* in, int W, int H, int R, int use_fma, float* out) {
    const size_t n = (size_t)W * H;
#pragma omp parallel for collapse(2) schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const uint8_t* p = img + 4 * ((size_t)y * W + x);
            float num = 0.00001f, den = 0.00001f;
            for (int i = 0; i < 2 * R + 1; i++) {
                int xx = clampi(x + i - R, 0, W - 1);
                size_t qi = (size_t)y * W + xx;
                float ww = supp(p, img + 4 * qi, abs(x - xx));               /* :41 */
                float F = conf[qi];                                           /* :42 */
                float wf = ww * F;
                float a = wf * in[qi];                                        /* :44  ww*F*v0*v1, left to right */
                if (use_fma) { num = fmaf(a, in[qi + n], num); den = fmaf(wf, in[qi + n], den); }
                else { num = num + a * in[qi + n]; den = den + wf * in[qi + n]; }   /* :44-45 */
            }
            out[(size_t)y * W + x] = num / den;                               /* :51 */
            out[(size_t)y * W + x + n] = den;                                 /* :52 */
        }
}

/* asw_WTA_REF -- kernels/asw_wta_ref.cl:2-68.  cost = final aggregated volume (x + W*y + W*H*d). */
ORACLE_API void oracle_asw_wta_ref(const float* cost, const float* ref, const float* ref_t, int W, int H, int D, int use_fma,
                                   uint8_t* out_rgba, uint8_t* out_tar_rgba, float* disp_ref, float* disp_ref_t, float* confidence) {
    const size_t n = (size_t)W * H;
    const float scale = (float)(D - 1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x;
            float cur = 100000.0f, last = 100000.0f;
            int min_d = 0;
            const float a = 0.085f * ref[p + n];                              /* :26 */
            for (int i = 0; i < D; i++) {
                float ad = fabsf(ref[p] - (float)i);
                float pen = use_fma ? fmaf(a, ad, cost[p + n * i]) : a * ad + cost[p + n * i];
                last = pen < last ? pen : last;                               /* :29-32 */
                min_d = pen < cur ? i : min_d;
                last = pen < cur ? cur : last;
                cur = pen < cur ? pen : cur;
            }
            int min_d_r = min_d;
            float cur_t = 100000.0f, last_t = 100000.0f;
            const float at = 0.085f * ref_t[p + n];
            for (int i = 0; i < min_d; i++) {
                int xq = x - i > 0 ? x - i : 0;
                int b = xq - x + min_d;                                       /* bresenham, asw_wta.cl:3-9 */
                float ad = fabsf(ref_t[p] - (float)i);                        /* QUIRK :46: the loop index i, not b */
                float c = cost[(size_t)xq + (size_t)W * y + n * b];
                float pen = use_fma ? fmaf(at, ad, c) : at * ad + c;
                last_t = pen < last_t ? pen : last_t;
                min_d_r = pen < cur_t ? b : min_d_r;
                last_t = pen < cur_t ? cur_t : last_t;
                cur_t = pen < cur_t ? pen : cur_t;
            }
            uint8_t v = D > 1 ? oracle_q8((float)min_d / scale) : 0, vt = D > 1 ? oracle_q8((float)min_d_r / scale) : 0;
            out_rgba[4 * p] = out_rgba[4 * p + 1] = out_rgba[4 * p + 2] = v; out_rgba[4 * p + 3] = 255;
            out_tar_rgba[4 * p] = out_tar_rgba[4 * p + 1] = out_tar_rgba[4 * p + 2] = vt; out_tar_rgba[4 * p + 3] = 255;
            if (disp_ref) disp_ref[p] = (float)min_d;
            if (disp_ref_t) disp_ref_t[p] = (float)min_d_r;
            /* QUIRK :63,66: both confidences are written to the SAME buffer, the target one last;
             * confidence_target is never rewritten */
            confidence[p] = (last_t - cur_t) / last_t;
        }
}

/* Median -- kernels/median.cl:58-88: per-channel 3x3 median, clamp-to-edge reads */
ORACLE_API void oracle_median(const uint8_t* in, int W, int H, uint8_t* out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int c = 0; c < 4; c++) {
                uint8_t s[9];
                int k = 0;
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++)
                        s[k++] = in[4 * ((size_t)clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)) + c];
                for (int i = 1; i < 9; i++) {   /* insertion sort; the min/max network of median.cl yields the same median */
                    uint8_t v = s[i];
                    int j = i - 1;
                    while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; j--; }
                    s[j + 1] = v;
                }
                out[4 * ((size_t)y * W + x) + c] = s[4];
            }
}

/* Whole ASW method: hot path (main.cpp:463-526) + consistency, k refinement rounds, median
 * (main.cpp:529-631).  Outputs (any may be NULL): final disparity image (asw_disparity.png),
 * the two consistency images (asw_consistency_pre-reff.png / post-reff.png). */
ORACLE_API int oracle_asw_full(const uint8_t* left, const uint8_t* right, int W, int H, const oracle_params* prm, int use_fma,
                               int refine_iters, uint8_t* out_disparity, uint8_t* out_pre_red, uint8_t* out_post_red) {
    const size_t n = (size_t)W * H;
    const int D = prm->ndisp, R = prm->radius;
    const float dscale = (float)(D - 1);
    float* cost = (float*)malloc(sizeof(float) * n * D);
    uint8_t *lw = (uint8_t*)malloc(4 * n), *rw = (uint8_t*)malloc(4 * n), *ce = (uint8_t*)malloc(4 * n), *red = (uint8_t*)malloc(4 * n);
    float *cr = (float*)malloc(sizeof(float) * n), *ct = (float*)malloc(sizeof(float) * n);
    float *vl = (float*)malloc(sizeof(float) * 2 * n), *vr = (float*)malloc(sizeof(float) * 2 * n);
    float *hl = (float*)malloc(sizeof(float) * 2 * n), *hr = (float*)malloc(sizeof(float) * 2 * n);
    int rc = -1;
    if (cost && lw && rw && ce && red && cr && ct && vl && vr && hl && hr &&
        oracle_asw_hot_path(left, right, W, H, prm, use_fma, cost, lw, rw, NULL, NULL, cr, ct) == 0) {
        oracle_consistency(lw, rw, W, H, dscale, cr, ct, ce, red);                           /* main.cpp:531-536 */
        if (out_pre_red) memcpy(out_pre_red, red, 4 * n);
        for (int i = 0; i < refine_iters; i++) {                                             /* main.cpp:545-614 */
            oracle_asw_ref_v(left, ce, cr, W, H, R, dscale, use_fma, vl);                    /* :547-552 (left estimate = consistency output) */
            oracle_asw_ref_v(right, rw, ct, W, H, R, dscale, use_fma, vr);                   /* :555-560 */
            oracle_asw_ref_h(left, cr, vl, W, H, R, use_fma, hl);                            /* :563-568 */
            oracle_asw_ref_h(right, ct, vr, W, H, R, use_fma, hr);                           /* :571-576 */
            oracle_asw_wta_ref(cost, hl, hr, W, H, D, use_fma, lw, rw, NULL, NULL, cr);      /* :579-589 */
            oracle_consistency(lw, rw, W, H, dscale, cr, ct, ce, red);                       /* :601-608 */
        }
        if (out_post_red) memcpy(out_post_red, red, 4 * n);
        if (out_disparity) oracle_median(ce, W, H, out_disparity);                           /* :617-619 */
        rc = 0;
    }
    free(cost); free(lw); free(rw); free(ce); free(red); free(cr); free(ct); free(vl); free(vr); free(hl); free(hr);
    return rc;
}
