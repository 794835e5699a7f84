/*
 * cross_oracle.c -- CPU restatement of the reference's second method, "Cross-Based Local Stereo
 * Matching Using Orthogonal Integral Images" (SURVEY.md section 8f, rank 4): median prefilter, cross
 * construction, raw cost, horizontal / vertical integral images with cross-limited box means, initial
 * WTA, cross-region voting, final median.  Host order: main.cpp:258-367.
 *
 * TEST INFRASTRUCTURE ONLY (see the header of asw_oracle.c).  Parity status: PINNED against the
 * reference's committed <dataset>/cross_based_initial.png and cross_based_disparity.png
 * (tests/test_oracle_golden.py).  Quirks of the reference kernels that the PNGs depend on are kept
 * and marked QUIRK.  Paths are relative to /root/reference/stereo_matching/.
 * The number of disparities (61) and the maximum arm length (25) are literals in the reference
 * (aggregation.cl:14, init_disparity.cl:11, disparity.cl:16, cross.cl:33-81); here they are arguments
 * whose defaults are those literals.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

uint8_t oracle_q8(float f);                                        /* asw_oracle.c: write_imagef UNORM8 rounding */
void oracle_median(const uint8_t* in, int W, int H, uint8_t* out); /* asw_tail_oracle.c: kernels/median.cl */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float unorm(uint8_t v) { return (float)v / 255.0f; }                       /* read_imagef, CL_UNORM_INT8 */
static inline const uint8_t* px_clamped(const uint8_t* img, int W, int H, int x, int y) { /* CLAMP_TO_EDGE sampler, main.cpp:10 */
    return img + 4 * ((size_t)clampi(y, 0, H - 1) * W + clampi(x, 0, W - 1));
}

/* check_similarity -- kernels/cross.cl:1-23 */
static int check_similarity(const uint8_t* img, int W, int H, int nx, int ny, int old_one, int current_one, const float* color) {
    const uint8_t* n = px_clamped(img, W, H, nx, ny);
    float check = 0.0f;
    for (int c = 0; c < 3; c++) check += (fabsf(color[c] - unorm(n[c])) < 0.10f) ? 1.0f : 0.0f;   /* :5-13 */
    int flag = ((float)(current_one - old_one) > 1.0f) ? 1 : 0;                                    /* :15 a gap ends the arm */
    current_one = (3.0f <= check) ? current_one : old_one;                                         /* :16 */
    flag += (nx < 0) + (ny < 0) + (W <= nx) + (H <= ny);                                           /* :17-20 */
    return flag ? old_one : current_one;                                                           /* :22 */
}

/* check_all -- kernels/cross.cl:25-81.  QUIRK: arm length k is tested on the pixel at distance k + 1
 * (the walk starts at offset + offset, :29). */
static int check_all(const uint8_t* img, int W, int H, int x, int y, int ox, int oy, int max_arm) {
    const uint8_t* p = px_clamped(img, W, H, x, y);
    const float color[3] = {unorm(p[0]), unorm(p[1]), unorm(p[2])};
    int arm = 1;
    for (int k = 1; k <= max_arm; k++) arm = check_similarity(img, W, H, x + (k + 1) * ox, y + (k + 1) * oy, arm, k, color);
    return arm;
}

/* Cross -- kernels/cross.cl:83-105.  out: 4 planes of W*H ints: -h_minus, h_plus, -v_minus, v_plus. */
ORACLE_API void oracle_cb_cross(const uint8_t* img, int W, int H, int max_arm, int32_t* out) {
    const size_t n = (size_t)W * H;
#pragma omp parallel for collapse(2) schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x;
            out[p] = -check_all(img, W, H, x, y, -1, 0, max_arm);
            out[p + n] = check_all(img, W, H, x, y, 1, 0, max_arm);
            out[p + 2 * n] = -check_all(img, W, H, x, y, 0, -1, max_arm);
            out[p + 3 * n] = check_all(img, W, H, x, y, 0, 1, max_arm);
        }
}

/* Aggregation -- kernels/aggregation.cl:3-23: SAD of the UNORM (0..1) colours, right pixel at x - d clamped */
ORACLE_API void oracle_cb_aggregation(const uint8_t* left, const uint8_t* right, int W, int H, int D, float* cost) {
    const size_t n = (size_t)W * H;
#pragma omp parallel for collapse(2) schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const uint8_t* l = left + 4 * ((size_t)y * W + x);
            for (int d = 0; d < D; d++) {
                const uint8_t* r = px_clamped(right, W, H, x - d, y);
                cost[(size_t)y * W + x + n * d] =
                    fabsf(unorm(l[0]) - unorm(r[0])) + fabsf(unorm(l[1]) - unorm(r[1])) + fabsf(unorm(l[2]) - unorm(r[2]));   /* :19 */
            }
        }
}

/* Integral_h -- kernels/integral_h.cl:3-17: in-place running sum along x, strictly left to right */
ORACLE_API void oracle_cb_integral_h(float* cost, int W, int H, int D) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < H * D; r++) {
        float* row = cost + (size_t)r * W;
        float sum = 0;
        for (int i = 0; i < W; i++) { sum = sum + row[i]; row[i] = sum; }
    }
}

/* Integral_v -- kernels/integral_v.cl:3-17: in-place running sum along y, strictly top to bottom */
ORACLE_API void oracle_cb_integral_v(float* cost, int W, int H, int D) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < D; d++)
        for (int x = 0; x < W; x++) {
            float* col = cost + (size_t)d * W * H + x;
            float sum = 0;
            for (int i = 0; i < H; i++) { sum = sum + col[(size_t)i * W]; col[(size_t)i * W] = sum; }
        }
}

/* Oii_hcross -- kernels/oii_hcross.cl:1-31: mean over the intersection of the left arm at x and the right
 * arm at max(0, x - d).  QUIRK: the window sum is I[x + h+] - I[x + h- - 1] but the divisor is h+ - h-
 * (one less than the number of summed pixels). */
ORACLE_API void oracle_cb_oii_hcross(const int32_t* cross_l, const int32_t* cross_r, const float* cost, int W, int H, int D, float* out) {
    const size_t n = (size_t)W * H;
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < D; d++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const size_t pr = (size_t)(x - d > 0 ? x - d : 0) + (size_t)y * W, pl = (size_t)x + (size_t)y * W;
                const int h_minus = cross_r[pr] > cross_l[pl] ? cross_r[pr] : cross_l[pl];                 /* :22 */
                const int h_plus = cross_r[pr + n] < cross_l[pl + n] ? cross_r[pr + n] : cross_l[pl + n]; /* :23 */
                const int delta = h_plus - h_minus;
                const float* row = cost + (size_t)y * W + n * d;
                const int hi = x + h_plus < W - 1 ? x + h_plus : W - 1, lo = x + h_minus - 1 > 0 ? x + h_minus - 1 : 0;
                out[pl + n * d] = (row[hi] - row[lo]) / (float)delta;                                      /* :26 */
            }
}

/* Oii_vcross -- kernels/oii_vcross.cl:1-32 */
ORACLE_API void oracle_cb_oii_vcross(const int32_t* cross_l, const int32_t* cross_r, const float* tcost, int W, int H, int D, float* out) {
    const size_t n = (size_t)W * H;
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < D; d++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const size_t pr = (size_t)(x - d > 0 ? x - d : 0) + (size_t)y * W, pl = (size_t)x + (size_t)y * W;
                const int v_minus = cross_l[pl + 2 * n] > cross_r[pr + 2 * n] ? cross_l[pl + 2 * n] : cross_r[pr + 2 * n];
                const int v_plus = cross_l[pl + 3 * n] < cross_r[pr + 3 * n] ? cross_l[pl + 3 * n] : cross_r[pr + 3 * n];
                const int delta = v_plus - v_minus;
                const float* col = tcost + x + n * d;
                const int hi = y + v_plus < H - 1 ? y + v_plus : H - 1, lo = y + v_minus - 1 > 0 ? y + v_minus - 1 : 0;
                out[pl + n * d] = (col[(size_t)hi * W] - col[(size_t)lo * W]) / (float)delta;             /* :26 */
            }
}

/* Init_disparity -- kernels/init_disparity.cl:1-19: strict-less argmin, lowest d wins ties */
ORACLE_API void oracle_cb_init_disparity(const float* cost, int W, int H, int D, uint8_t* out_rgba) {
    const size_t n = (size_t)W * H;
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; p++) {
        int min_d = 0;
        float min_result = cost[p];
        for (int i = 0; i < D; i++) {
            const float c = cost[p + n * i];
            if (c < min_result) { min_d = i; min_result = c; }
        }
        const uint8_t v = D > 1 ? oracle_q8((float)min_d / (float)(D - 1)) : 0;                            /* :17 */
        out_rgba[4 * p] = out_rgba[4 * p + 1] = out_rgba[4 * p + 2] = v;
        out_rgba[4 * p + 3] = 255;
    }
}

/* Disparity -- kernels/disparity.cl:1-41: histogram of the initial disparities over the pixel's cross
 * region (vertical arm of the pixel; per row the horizontal arm of the pixel (x, y + i)), most frequent
 * value wins.  QUIRKs: the bin is (int)(v/255 * (D-1)) truncated (:29-30); ties go to the LARGER
 * disparity (:35-36 keep the old index only when strictly less). */
ORACLE_API void oracle_cb_disparity(const uint8_t* init_rgba, const int32_t* cross, int W, int H, int D, uint8_t* out_rgba) {
    const size_t n = (size_t)W * H;
    const float scale = (float)(D - 1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x;
            int tab[256];
            for (int i = 0; i < D; i++) tab[i] = 0;
            const int v_minus = cross[p + 2 * n], v_plus = cross[p + 3 * n];
            for (int i = v_minus; i <= v_plus; i++) {
                const size_t q = (size_t)x + (size_t)clampi(y + i, 0, H - 1) * W;
                const int h_minus = cross[q], h_plus = cross[q + n];
                for (int j = h_minus; j <= h_plus; j++) {
                    const float v = unorm(px_clamped(init_rgba, W, H, x + j, y + i)[0]) * scale;          /* :29 */
                    tab[(int)v]++;                                                                         /* :30 */
                }
            }
            int result = 0, result_indx = 0;
            for (int i = 0; i < D; i++)
                if (!((float)tab[i] < (float)result)) { result_indx = i; result = tab[i]; }               /* :35-36 */
            const uint8_t v = D > 1 ? oracle_q8((float)((double)result_indx / (double)(D - 1))) : 0;      /* :38 (60.0 is a double literal) */
            out_rgba[4 * p] = out_rgba[4 * p + 1] = out_rgba[4 * p + 2] = v;
            out_rgba[4 * p + 3] = 255;
        }
}

/* Median as the reference's host launches it for this method: NDRange = local * floor(dim / local) with
 * local = {3, 3} (main.cpp:191,193,197,274,279,354 -- `ceil` of an integer division), so the last W % 3 columns
 * and H % 3 rows are never written; the freshly created images read back as zeros there (QUIRK, visible
 * in art/cross_based_disparity.png: alpha = 0 in its last two rows).  local = 1 covers the whole image. */
ORACLE_API void oracle_cb_median_grid(const uint8_t* in, int W, int H, int local, uint8_t* out) {
    oracle_median(in, W, H, out);
    if (local > 1) {
        const int We = local * (W / local), He = local * (H / local);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                if (x >= We || y >= He) memset(out + 4 * ((size_t)y * W + x), 0, 4);
    }
}

/* The whole method, main.cpp:270-367.  Outputs (any may be NULL): cross_based_initial.png (`disparity`),
 * cross_based_disparity.png (median of the voted map), median.png (median of the left image). */
ORACLE_API int oracle_cross_full(const uint8_t* left, const uint8_t* right, int W, int H, int D, int max_arm, int median_local, uint8_t* out_initial,
                                 uint8_t* out_final, uint8_t* out_median_l) {
    const size_t n = (size_t)W * H;
    if (!left || !right || W <= 0 || H <= 0 || D <= 0 || D > 256 || max_arm < 1) return -1;
    uint8_t *ml = (uint8_t*)malloc(4 * n), *mr = (uint8_t*)malloc(4 * n), *init = (uint8_t*)malloc(4 * n), *voted = (uint8_t*)malloc(4 * n);
    int32_t *cl = (int32_t*)malloc(16 * n), *cr = (int32_t*)malloc(16 * n);
    float *cost = (float*)malloc(sizeof(float) * n * D), *tmp = (float*)malloc(sizeof(float) * n * D);
    int rc = -1;
    if (ml && mr && init && voted && cl && cr && cost && tmp) {
        oracle_cb_median_grid(left, W, H, median_local, ml);             /* main.cpp:270-279 */
        oracle_cb_median_grid(right, W, H, median_local, mr);
        oracle_cb_cross(ml, W, H, max_arm, cl);                          /* :283-291 */
        oracle_cb_cross(mr, W, H, max_arm, cr);
        oracle_cb_aggregation(ml, mr, W, H, D, cost);                    /* :295-299 */
        oracle_cb_integral_h(cost, W, H, D);                             /* :304-307 */
        oracle_cb_oii_hcross(cl, cr, cost, W, H, D, tmp);                /* :312-318 */
        oracle_cb_integral_v(tmp, W, H, D);                              /* :322-325 */
        oracle_cb_oii_vcross(cl, cr, tmp, W, H, D, cost);                /* :329-335 */
        oracle_cb_init_disparity(cost, W, H, D, init);                   /* :339-342 */
        oracle_cb_disparity(init, cl, W, H, D, voted);                   /* :346-350 */
        if (out_initial) memcpy(out_initial, init, 4 * n);               /* :357-359 */
        if (out_final) oracle_cb_median_grid(voted, W, H, median_local, out_final);   /* :352-354, 361-363 */
        if (out_median_l) memcpy(out_median_l, ml, 4 * n);               /* :365-367 */
        rc = 0;
    }
    free(ml); free(mr); free(init); free(voted); free(cl); free(cr); free(cost); free(tmp);
    return rc;
}
