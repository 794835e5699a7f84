/*
 * asw_b200.h -- C ABI of the B200-native ASW (adaptive support weight) stereo hot path.
 *
 * This is the drop-in boundary for ONE path of manixq/stereo_matchin: the OpenCL ASW
 * pipeline of stereo_matching/kernels up to winner-take-all
 *   raw SAD cost -> 4 support-weight tables -> r x (vertical, horizontal) joint-weight
 *   aggregation -> WTA
 * that the reference drives from stereo_matching/main.cpp:412-526.  The reference has no
 * plugin API: its operator surface is the six `__kernel` signatures plus the host
 * sequence in main.cpp.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference's stereo_matching/ directory).
 *
 * Conventions (identical to the reference so buffers are interchangeable with it):
 *   - images: RGBA8, tightly packed rows (pitch W*4), origin top-left
 *     (main.cpp:189 `format = {CL_RGBA, CL_UNORM_INT8}`, main.cpp:243-244); the kernels read
 *     them through a CLAMP_TO_EDGE / nearest sampler (main.cpp:10) and ignore alpha.
 *   - cost volumes: float32, index x + W*y + W*H*d            (kernels/asw_aggr.cl:21)
 *   - support tables: float32, index x + W*y + W*H*tap, tap = 0..2*radius
 *                                                              (kernels/asw_vsupport.cl:26)
 *   - disparity images: RGBA8, grey replicated, A = 255, grey = q8(d / (ndisp-1))
 *                                                              (kernels/asw_wta.cl:70,73)
 * All entry points return 0 (ASW_OK) or an asw_status code; nothing throws or exits
 * across the ABI (the reference only prints on error, main.cpp:27-30).
 * A context is bound to one GPU and one CUDA stream and is NOT thread-safe: use one
 * context per GPU / host thread (the reference uses one in-order queue, main.cpp:212).
 * There is no CPU fallback: without a CUDA device asw_create fails.
 */
#ifndef ASW_B200_H
#define ASW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ASW_API __attribute__((visibility("default")))
#else
#define ASW_API
#endif

typedef struct asw_ctx asw_ctx;

typedef enum asw_status {
    ASW_OK = 0,
    ASW_ERR_INVALID = 1,      /* NULL pointer / non-positive size / bad parameter */
    ASW_ERR_CUDA = 2,         /* a CUDA runtime call failed; see asw_last_error() */
    ASW_ERR_NOMEM = 3,        /* device or host allocation failed */
    ASW_ERR_UNSUPPORTED = 4   /* parameter outside the implemented range */
} asw_status;

/* Algorithm parameters.  The reference hard-codes all of them as literals; the defaults
 * written by asw_params_default() reproduce those literals exactly. */
typedef struct asw_params {
    int radius;       /* support radius R, taps T = 2R+1.  16  (asw_vsupport.cl:19, asw_vcost_aggregation.cl:33,36) */
    int ndisp;        /* number of disparities D (0..D-1). 61  (asw_aggr.cl:16, asw_wta.cl:34) */
    float gamma_c;    /* colour bandwidth.                 30.91f (asw_vsupport.cl:22) */
    float gamma_p;    /* proximity bandwidth.              28.21f (asw_vsupport.cl:24) */
    float trunc;      /* raw-cost truncation; +inf = none (the reference does not truncate, asw_aggr.cl:19) */
    int iterations;   /* r, number of (V,H) aggregation rounds. 7 (main.cpp:177) */
} asw_params;

/* Per-stage device times in milliseconds (CUDA events on the context's stream); the
 * same columns the reference writes to its per-device log from OpenCL events
 * (main.cpp:634-660: aggr, supp_w, v_aggr_mean, h_aggr_mean, total aggregation, wta).
 * Stages that are fused into a neighbour report 0 and fold into agg_total_ms. */
typedef struct asw_timing {
    float raw_ms;         /* asw_Aggr                              (main.cpp:634) */
    float supp_ms;        /* the four support-table launches       (main.cpp:635) */
    float vagg_mean_ms;   /* mean vertical pass                    (main.cpp:640-644) */
    float hagg_mean_ms;   /* mean horizontal pass                  (main.cpp:650-656) */
    float agg_total_ms;   /* first V start .. last H end           (main.cpp:658-659) */
    float wta_ms;         /* asw_WTA                               (main.cpp:660) */
    float total_ms;       /* raw start .. WTA end (hot path, device only) */
    float h2d_ms;         /* host->device upload (host-buffer entry points only) */
    float d2h_ms;         /* device->host download */
    int kernel_launches;  /* CUDA kernels launched by the call */
    float vfix_mean_ms;   /* part of vagg_mean_ms spent in the small fix-up / edge-padding launches that follow the
                             main vertical kernel (0 for kernel families that have none) */
} asw_timing;

/* Device times of the consumers of the hot path, the remaining columns of the reference's log
 * (main.cpp:661-708: consistency, v_ref_mean_L/R, h_ref_mean_L/R, wta_mean_LR, consistency_mean,
 * total refinement, median, total WTA method).  Filled by asw_stereo. */
typedef struct asw_tail_timing {
    float right_wta_ms;          /* right-view part of asw_WTA (the reference computes it inside asw_WTA) */
    float consistency_ms;        /* first Constistency                              (main.cpp:531-536) */
    float vref_mean_l_ms;        /* mean asw_ref_v, left view                       (main.cpp:547-552) */
    float vref_mean_r_ms;        /* mean asw_ref_v, right view                      (main.cpp:555-560) */
    float href_mean_l_ms;        /* mean asw_ref_h, left view                       (main.cpp:563-568) */
    float href_mean_r_ms;        /* mean asw_ref_h, right view                      (main.cpp:571-576) */
    float wta_ref_mean_ms;       /* mean asw_WTA_REF                                (main.cpp:579-589) */
    float consistency_mean_ms;   /* mean Constistency inside the refinement loop    (main.cpp:601-608) */
    float refinement_total_ms;   /* all refinement rounds */
    float median_ms;             /* Median                                          (main.cpp:617-619) */
    float total_ms;              /* hot path + tail, device only ("total WTA method") */
} asw_tail_timing;

ASW_API const char* asw_version(void);
ASW_API const char* asw_strerror(int status);

/* Replaces the OpenCL platform/device/context/program/queue setup, main.cpp:119-130,158-172,210-212.
 * `device` is a CUDA ordinal. */
ASW_API int asw_create(asw_ctx** out, int device);
ASW_API int asw_destroy(asw_ctx* ctx);
ASW_API const char* asw_last_error(asw_ctx* ctx);
/* The context's cudaStream_t (as void*), so callers can order their own work after ours. */
ASW_API void* asw_stream(asw_ctx* ctx);
ASW_API int asw_sync(asw_ctx* ctx);
ASW_API int asw_device_info(asw_ctx* ctx, int* sm_count, int* sm_clock_khz, size_t* total_mem, char* name, size_t name_len);

ASW_API void asw_params_default(asw_params* p);

/* ---- fused hot path ---------------------------------------------------------------
 * Replaces the enqueue sequence main.cpp:463-526 (asw_Aggr, 4x support, r x (V,H), asw_WTA)
 * plus the upload at main.cpp:243-244 and the blocking read at main.cpp:621.
 * Host buffers in, host buffers out; synchronous on return.  Any output may be NULL.
 *   disp_rgba : W*H*4  left disparity image exactly as asw_WTA's `output` image
 *   disp_d    : W*H    raw winning disparity index (uint8 when ndisp <= 256)
 *   conf      : W*H    (min2-min1)/min2, asw_WTA's confidence_reference */
ASW_API int asw_disparity(asw_ctx* ctx, const uint8_t* left_rgba, const uint8_t* right_rgba, int W, int H,
                          const asw_params* prm, uint8_t* disp_rgba, uint8_t* disp_d, float* conf, asw_timing* timing);

/* The same call without the final wait: upload, hot path and download are enqueued on the context's stream and the
 * call returns; asw_sync(ctx) completes them.  Host buffers must be pinned (asw_host_alloc) for the copies to overlap
 * other work, and must stay valid until the sync.  Two contexts on one device alternate between pairs of a batch
 * (cfg5: 1024 pairs) so that the upload of pair i+1 runs under the kernels of pair i -- the reference uploads at buffer
 * creation and reads back blocking (main.cpp:243-244,621), i.e. it has no such overlap. */
ASW_API int asw_disparity_async(asw_ctx* ctx, const uint8_t* left_rgba, const uint8_t* right_rgba, int W, int H,
                                const asw_params* prm, uint8_t* disp_rgba, uint8_t* disp_d, float* conf);

/* Same with DEVICE pointers for inputs and outputs; asynchronous on the context's stream
 * unless `timing` is non-NULL (timing needs the events to complete).
 * Launch structure: when a call repeats an earlier one's signature (same buffers, shape, band and parameters; the
 * CONTENTS of the buffers may differ -- the next frame in the same buffers), its ~26 kernel launches are captured into a
 * CUDA graph on the second occurrence and replayed from the third on (up to 4 signatures per context; any growth of the
 * context's scratch memory drops them; calls with `timing`, a kept volume, a shard or a halo exchange are never
 * captured).  The reference re-enqueues its kernels one by one for every frame (main.cpp:463-526). */
ASW_API int asw_disparity_device(asw_ctx* ctx, const uint8_t* d_left_rgba, const uint8_t* d_right_rgba, int W, int H,
                                 const asw_params* prm, uint8_t* d_disp_rgba, uint8_t* d_disp_d, float* d_conf,
                                 asw_timing* timing);

/* Row-band variant for multi-GPU sharding: computes output rows [y0, y1) of the W x H
 * frame only.  Inputs are the FULL images (device pointers); outputs hold (y1-y0) rows.
 * Internally rows [y0 - r*R, y1 + r*R) are processed with a halo that shrinks by R per
 * iteration, so the band is bit-identical to the same rows of asw_disparity_device. */
ASW_API int asw_disparity_band_device(asw_ctx* ctx, const uint8_t* d_left_rgba, const uint8_t* d_right_rgba, int W, int H,
                                      int y0, int y1, const asw_params* prm, uint8_t* d_disp_rgba, uint8_t* d_disp_d,
                                      float* d_conf, asw_timing* timing);

/* Row-band variant with a per-iteration halo exchange (alternative (i)): the band keeps only `radius`
 * halo rows and every iteration aggregates exactly the rows [y0, y1) -- no halo rows are recomputed.  After the
 * horizontal pass of every iteration but the last, `exchange` is called with DEVICE pointers into the band's cost
 * volume: `top_send` / `bottom_send` are this band's first / last `radius` rows (`bytes` each, contiguous), to be
 * delivered to the `bottom_recv` of the band above / the `top_recv` of the band below; `top_recv` / `bottom_recv` is
 * where the neighbours' rows belong.  Pointers towards a frame border are NULL.  The work of the iteration has been
 * ENQUEUED on asw_stream(ctx) when the callback runs: it must order its transfers after that work and must not return
 * before this band's receive buffers are complete and its send buffers may be overwritten (every band's next
 * horizontal pass rewrites its send rows).  Returns non-zero to abort.  The result is bit-identical to the same rows of
 * asw_disparity_device: the reference itself has no multi-device split (main.cpp:158-172 visits devices one by one). */
typedef int (*asw_halo_fn)(void* user, int iteration, void* top_send, void* bottom_send, void* top_recv, void* bottom_recv,
                           size_t bytes);
ASW_API int asw_disparity_band_exchange_device(asw_ctx* ctx, const uint8_t* d_left_rgba, const uint8_t* d_right_rgba, int W,
                                               int H, int y0, int y1, const asw_params* prm, uint8_t* d_disp_rgba,
                                               uint8_t* d_disp_d, float* d_conf, asw_halo_fn exchange, void* user,
                                               asw_timing* timing);

/* The same exchange hidden under the computation: no call blocks the host.  Every iteration computes the `radius`
 * boundary rows of its horizontal pass first, on a second stream beside the interior rows, and calls `begin` with the
 * same pointers as above plus that stream (a cudaStream_t): the transfer must be ordered after the work enqueued on
 * it and is left running.  The next iteration aggregates the interior rows of its vertical pass -- they read no halo
 * row -- and then calls `end` with the main stream (asw_stream): the callback makes that stream wait for the transfer
 * (receive buffers complete, send buffers free), after which the band's border rows are aggregated.  Neither callback
 * may wait on the host for another band's GPU work. */
typedef int (*asw_halo_begin_fn)(void* user, int iteration, void* top_send, void* bottom_send, void* top_recv, void* bottom_recv,
                                 size_t bytes, void* boundary_stream);
typedef int (*asw_halo_end_fn)(void* user, int iteration, void* main_stream);
ASW_API int asw_disparity_band_exchange_async_device(asw_ctx* ctx, const uint8_t* d_left_rgba, const uint8_t* d_right_rgba,
                                                     int W, int H, int y0, int y1, const asw_params* prm, uint8_t* d_disp_rgba,
                                                     uint8_t* d_disp_d, float* d_conf, asw_halo_begin_fn begin,
                                                     asw_halo_end_fn end, void* user, asw_timing* timing);
/* the second stream of the context (boundary rows of the horizontal pass), as void* */
ASW_API void* asw_side_stream(asw_ctx* ctx);

/* ---- one frame on several GPUs of one process -----------------------------------------------------
 * Replaces the reference's device loop (main.cpp:119-130,158-172, which runs the WHOLE job on every device in turn)
 * by a split of one frame: row bands, one band per listed CUDA device, `radius` boundary rows pulled from the
 * neighbouring bands' volumes over NVLink between iterations (cudaMemcpyPeerAsync; no rows are recomputed).  Host
 * buffers in and out as asw_disparity; the result is bit-identical to asw_disparity on one device. */
typedef struct asw_multi asw_multi;
typedef struct asw_multi_timing {
    int devices;
    float upload_ms;                /* host wall clock: image rows to every device */
    float compute_ms;               /* host wall clock: first launch to the slowest band's last kernel */
    float download_ms;              /* host wall clock: bands of the result into the caller's buffers */
    float total_ms;
    float slowest_band_device_ms;   /* CUDA-event time of the slowest band (includes its waits for the neighbours) */
} asw_multi_timing;
ASW_API int asw_multi_create(asw_multi** out, const int* devices, int n_devices);
ASW_API int asw_multi_destroy(asw_multi* m);
ASW_API int asw_multi_count(asw_multi* m);
ASW_API const char* asw_multi_last_error(asw_multi* m);
ASW_API int asw_multi_disparity(asw_multi* m, const uint8_t* left_rgba, const uint8_t* right_rgba, int W, int H,
                                const asw_params* prm, uint8_t* disp_rgba, uint8_t* disp_d, float* conf,
                                asw_multi_timing* timing);

/* Disparity-shard variant for multi-GPU sharding without halo work: aggregates only the disparities [d0, d1)
 * (d0 a multiple of 64) for output rows [y0, y1) and returns the shard's partial winner-take-all result per pixel:
 * smallest and second smallest aggregated cost and the GLOBAL index of the smallest ((y1-y0)*W each).  Planes of the
 * cost volume never interact during aggregation (asw_vcost_aggregation.cl / asw_hcost_aggregation.cl: d is a pure
 * index), so shards are independent; asw_merge_shards combines them.  Needs the default kernel family, radius 16. */
ASW_API int asw_disparity_shard_device(asw_ctx* ctx, const uint8_t* d_left_rgba, const uint8_t* d_right_rgba, int W, int H,
                                       int y0, int y1, int d0, int d1, const asw_params* prm, float* d_min1, float* d_min2,
                                       int* d_arg, asw_timing* timing);
/* Combines `nshards` partial results ([shard][rows][W] arrays, shards in ascending disparity order, e.g. straight
 * out of an all-gather) into the outputs of asw_WTA; bit-identical to the unsharded call.  Any output may be NULL. */
ASW_API int asw_merge_shards(asw_ctx* ctx, int W, int rows, int ndisp, int nshards, const float* d_min1, const float* d_min2,
                             const int* d_arg, uint8_t* d_disp_rgba, uint8_t* d_disp_d, float* d_conf);

/* Keep the final aggregated volume of the last asw_disparity* call (device pointer, layout
 * above, rows of the processed band) for consumers such as the right-view WTA / consistency
 * check.  Returns NULL if the last call did not hand it out.  Enable with
 * asw_set_keep_volume(ctx, 1): the final volume is then converted from the kernels' own
 * layout to the reference layout (one extra pass over the volume).  Default 0.  With the
 * default 0 the last horizontal pass takes the winner inside its epilogue (per 128-disparity
 * window, merged by a small kernel) and the final volume is never written to HBM; with 1
 * (and for disparity shards) the volume is written and the separate WTA kernel reads it.
 * Both give identical bits. */
ASW_API int asw_set_keep_volume(asw_ctx* ctx, int keep);
ASW_API const float* asw_final_volume(asw_ctx* ctx);

/* ---- per-operator entry points (device pointers; reference layouts) ------------------
 * Argument order mirrors the clSetKernelArg sequences in main.cpp; OpenCL obtained W,H
 * from get_image_dim(), here they are explicit. */

/* kernels/asw_aggr.cl:3-23 `asw_Aggr(input_l, input_r, output_cost)`; args main.cpp:463-465 */
ASW_API int asw_Aggr(asw_ctx* ctx, const uint8_t* d_input_l, const uint8_t* d_input_r, int W, int H,
                     const asw_params* prm, float* d_output_cost);
/* kernels/asw_vsupport.cl:3-27 `asw_vSupport(input, output_cost)`; args main.cpp:470-471 */
ASW_API int asw_vSupport(asw_ctx* ctx, const uint8_t* d_input, int W, int H, const asw_params* prm, float* d_output);
/* kernels/asw_hsupport.cl:3-28 `asw_hSupport(input, output_cost)`; args main.cpp:474-475 */
ASW_API int asw_hSupport(asw_ctx* ctx, const uint8_t* d_input, int W, int H, const asw_params* prm, float* d_output);
/* kernels/asw_vcost_aggregation.cl:11-44
 * `asw_vCostAggregation(input_l, supp_left, supp_right, input_cost, output_denom, output_cost)`;
 * args main.cpp:494-499.  input_l was only used for its dimensions.  output_denom may be NULL. */
ASW_API int asw_vCostAggregation(asw_ctx* ctx, int W, int H, const asw_params* prm, const float* d_supp_left,
                                 const float* d_supp_right, const float* d_input_cost, float* d_output_denom,
                                 float* d_output_cost);
/* kernels/asw_hcost_aggregation.cl:12-44
 * `asw_hCostAggregation(input_l, supp_left, supp_right, vertical_cost, denom_v, output_cost)`;
 * args main.cpp:503-508.  denom_v is accepted and ignored, as in the reference (:17). */
ASW_API int asw_hCostAggregation(asw_ctx* ctx, int W, int H, const asw_params* prm, const float* d_supp_left,
                                 const float* d_supp_right, const float* d_vertical_cost, const float* d_denom_v,
                                 float* d_output_cost);
/* kernels/asw_wta.cl:12-82
 * `asw_WTA(output_cost, output, d_est_reference, d_est_target, output_target,
 *          confidence_reference, confidence_target)`; args main.cpp:519-525.
 * Any output may be NULL; the right/target outputs (asw_wta.cl:50-67) are computed only
 * if one of them is requested. */
ASW_API int asw_WTA(asw_ctx* ctx, int W, int H, const asw_params* prm, const float* d_cost, uint8_t* d_output_rgba,
                    float* d_est_reference, float* d_est_target, uint8_t* d_output_target_rgba,
                    float* d_confidence_reference, float* d_confidence_target);

/* ---- consumers of the hot path ("next" rows): consistency, refinement, penalised WTA, median ----
 * Device pointers, reference layouts; argument order mirrors the clSetKernelArg sequences. */

/* kernels/consist.cl:3-34 `Constistency(ref, tar, confidence_ref, confidence_tar, output, output_red)`;
 * args main.cpp:531-536.  Disparity images are read as v/255*(ndisp-1); output / output_red may be NULL. */
ASW_API int asw_Constistency(asw_ctx* ctx, int W, int H, const asw_params* prm, const uint8_t* d_ref_rgba,
                             const uint8_t* d_tar_rgba, float* d_confidence_ref, float* d_confidence_tar,
                             uint8_t* d_output_rgba, uint8_t* d_output_red_rgba);
/* kernels/asw_refinement_v.cl:13-51 `asw_ref_v(input, input_est, confidence, output_REF)`; args main.cpp:547-551.
 * output_REF holds two W*H planes (value, denominator). */
ASW_API int asw_ref_v(asw_ctx* ctx, int W, int H, const asw_params* prm, const uint8_t* d_input_rgba,
                      const uint8_t* d_input_est_rgba, const float* d_confidence, float* d_output_REF);
/* kernels/asw_refinement_h.cl:16-53 `asw_ref_h(input, confidence, input_REF, output_REF)`; args main.cpp:563-567 */
ASW_API int asw_ref_h(asw_ctx* ctx, int W, int H, const asw_params* prm, const uint8_t* d_input_rgba, const float* d_confidence,
                      const float* d_input_REF, float* d_output_REF);
/* kernels/asw_wta_ref.cl:2-68 `asw_WTA_REF(agg_d, ref, ref_target, output, output_target, disp_ref,
 * disp_ref_target, confidence, confidence_target)`; args main.cpp:580-588.  As in the reference, the
 * target confidence overwrites `confidence` and confidence_target is left untouched (:63,66). */
ASW_API int asw_WTA_REF(asw_ctx* ctx, int W, int H, const asw_params* prm, const float* d_agg_d, const float* d_ref,
                        const float* d_ref_target, uint8_t* d_output_rgba, uint8_t* d_output_target_rgba, float* d_disp_ref,
                        float* d_disp_ref_target, float* d_confidence, float* d_confidence_target);
/* kernels/median.cl:58-88 `Median(input, output)`; args main.cpp:617-618 */
ASW_API int asw_Median(asw_ctx* ctx, int W, int H, const uint8_t* d_input_rgba, uint8_t* d_output_rgba);

/* The whole ASW method of main.cpp:463-631 on host buffers: hot path (fused kernels), consistency,
 * `refine_iters` (reference: k = 6, main.cpp:176) refinement rounds, median.  Outputs (any may be
 * NULL), each W*H*4: the final disparity image (the reference's asw_disparity.png) and the two
 * consistency images (asw_consistency_pre-reff.png / asw_consistency_post-reff.png). */
ASW_API int asw_stereo(asw_ctx* ctx, const uint8_t* left_rgba, const uint8_t* right_rgba, int W, int H, const asw_params* prm,
                       int refine_iters, uint8_t* disparity_rgba, uint8_t* consistency_pre_rgba, uint8_t* consistency_post_rgba,
                       asw_timing* timing, asw_tail_timing* tail_timing);

/* ---- the reference's second method: cross-based matching on orthogonal integral images -------------
 * (main.cpp:258-367; kernels cross.cl, aggregation.cl, integral_h.cl, oii_hcross.cl, integral_v.cl,
 * oii_vcross.cl, init_disparity.cl, disparity.cl, median.cl).  Device pointers, reference layouts:
 * cost volumes x + W*y + W*H*d; cross tables 4 planes of W*H ints (-h_minus, h_plus, -v_minus, v_plus). */
typedef struct asw_cross_params {
    int ndisp;          /* 61   (aggregation.cl:14, init_disparity.cl:11, disparity.cl:16; at most 256) */
    int max_arm;        /* 25   (cross.cl:33-81) */
    int median_local;   /* 3: the reference launches Median on local * floor(dim / local) work items (main.cpp:191-197),
                           pixels beyond stay zero, as in its committed PNGs; 1 = filter the whole image */
} asw_cross_params;

/* the reference's log columns for this method (main.cpp:379-396), device times in ms */
typedef struct asw_cross_timing {
    float median_l_ms, median_r_ms, median_ms, cross_l_ms, cross_r_ms, cross_ms, aggregation_ms, integral_h_ms, oii_h_ms,
        integral_v_ms, oii_v_ms, init_disparity_ms, final_disparity_ms, total_ms;
} asw_cross_timing;

ASW_API void asw_cross_params_default(asw_cross_params* p);
/* Median as launched for this method: asw_Median, then the pixels outside local * floor(dim / local) zeroed */
ASW_API int asw_Median_grid(asw_ctx* ctx, int W, int H, int local, const uint8_t* d_input_rgba, uint8_t* d_output_rgba);
/* kernels/cross.cl:83-105 `Cross(input, output)`; args main.cpp:283-291 */
ASW_API int asw_Cross(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const uint8_t* d_input_rgba, int* d_output);
/* kernels/aggregation.cl:3-23 `Aggregation(input_l, input_r, output_cost)`; args main.cpp:295-298 */
ASW_API int asw_Aggregation(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const uint8_t* d_input_l_rgba,
                            const uint8_t* d_input_r_rgba, float* d_output_cost);
/* kernels/integral_h.cl:3-17, integral_v.cl:3-17 `Integral_h/v(cost, size)`: in place; args main.cpp:304-305, 322-323 */
ASW_API int asw_Integral_h(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, float* d_cost);
ASW_API int asw_Integral_v(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, float* d_cost);
/* kernels/oii_hcross.cl:1-31 `Oii_hcross(cross_l, cross_r, cost, temp_cost, size)`; args main.cpp:312-316 */
ASW_API int asw_Oii_hcross(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const int* d_cross_l, const int* d_cross_r,
                           const float* d_cost, float* d_temp_cost);
/* kernels/oii_vcross.cl:1-32 `Oii_vcross(cross_l, cross_r, temp_cost, cost, size)`; args main.cpp:329-333 */
ASW_API int asw_Oii_vcross(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const int* d_cross_l, const int* d_cross_r,
                           const float* d_temp_cost, float* d_cost);
/* kernels/init_disparity.cl:1-19 `Init_disparity(cost, output)`; args main.cpp:339-340 */
ASW_API int asw_Init_disparity(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const float* d_cost, uint8_t* d_output_rgba);
/* kernels/disparity.cl:1-41 `Disparity(input, input_cross, output)`; args main.cpp:346-348 */
ASW_API int asw_Disparity(asw_ctx* ctx, int W, int H, const asw_cross_params* prm, const uint8_t* d_input_rgba, const int* d_input_cross,
                          uint8_t* d_output_rgba);
/* The whole method on host buffers (main.cpp:270-367).  Outputs (any may be NULL), each W*H*4: the reference's
 * cross_based_initial.png, cross_based_disparity.png and median.png. */
ASW_API int asw_cross_stereo(asw_ctx* ctx, const uint8_t* left_rgba, const uint8_t* right_rgba, int W, int H,
                             const asw_cross_params* prm, uint8_t* initial_rgba, uint8_t* disparity_rgba, uint8_t* median_l_rgba,
                             asw_cross_timing* timing);

/* ---- device memory helpers (so a plain C/C++ host needs no CUDA headers) -------------
 * Replace clCreateBuffer / clCreateImage2D(COPY_HOST_PTR) / clEnqueueReadImage /
 * clReleaseMemObject, main.cpp:243-256,434-457,621-629,712-738. */
ASW_API int asw_dev_alloc(asw_ctx* ctx, void** d_ptr, size_t bytes);
ASW_API int asw_dev_free(asw_ctx* ctx, void* d_ptr);
ASW_API int asw_memcpy_h2d(asw_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
ASW_API int asw_memcpy_d2h(asw_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
/* pinned host staging memory (cudaHostAlloc) for overlapped uploads */
ASW_API int asw_host_alloc(asw_ctx* ctx, void** h_ptr, size_t bytes);
ASW_API int asw_host_free(asw_ctx* ctx, void* h_ptr);

/* Selects the kernel family of the fused path: 0 = automatic (default: the TMA-fed sm_100a kernels for
 * radius 16, ndisp padded to a multiple of 64 internally; the generic kernels for any other radius),
 * 1 = the generic one-thread-per-output kernels for every radius (an on-device cross-check of family 0).
 * Both are CUDA; there is no CPU path. */
ASW_API int asw_set_kernel_family(asw_ctx* ctx, int family);

#ifdef __cplusplus
}
#endif
#endif /* ASW_B200_H */
