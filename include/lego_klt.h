/*
 * lego_klt.h -- C ABI of the B200-native pyramid Gauss-Newton KLT tracker.
 *
 * This is the drop-in boundary for ONE hot path of LEGO-SLAM (reference paths are relative to the
 * upstream tree):
 *
 *   legoslam::LKOpticalFlow1Layer   include/legoslam/algorithm.h:123-128, src/algorithm.cpp:11-31
 *   legoslam::LKOpticalFlow4Layer   include/legoslam/algorithm.h:131-136, src/algorithm.cpp:128-206
 *   LKOpticalFlowTracker::calcLKOpticalFlow            src/algorithm.cpp:37-125   (the per-feature solver)
 *   GetPixelValue / IsPtInImg       include/legoslam/algorithm.h:40-66           (sampler / in-image test)
 *
 * called from Frontend::TrackLastFrameLKOpticalFlow4LayerSelf (src/frontend_g2o.cpp:473) and
 * Frontend::FindFeaturesInRightLKOpticalFlow4LayerSelf (src/frontend_g2o.cpp:515).
 *
 * Everything here is plain C: pointers, sizes, PODs.  No torch / OpenCV / Eigen types.  The C++
 * shim that keeps the reference's own signatures lives in include/legoslam_gpu/algorithm_shim.h.
 *
 * There is NO CPU fallback behind these entry points: every compute call runs hand-written sm_100a
 * CUDA kernels and returns a negative error code if no CUDA device / kernel image is available.
 */
#ifndef LEGO_KLT_H
#define LEGO_KLT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEGO_KLT_ABI_VERSION 1
#define LEGO_KLT_MAX_LEVELS 8

/* ---- error codes (reference returns void and has no status; see SURVEY.md 8b "Errors") ---- */
enum {
    LEGO_KLT_OK = 0,
    LEGO_KLT_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, levels out of range ...   */
    LEGO_KLT_ERR_CUDA = -2,         /* a CUDA runtime call failed; see lego_klt_last_error()       */
    LEGO_KLT_ERR_NO_DEVICE = -3,    /* no CUDA device: the product path never falls back to CPU    */
    LEGO_KLT_ERR_UNSUPPORTED = -4,  /* e.g. pyramid level would be empty, patch too large          */
    LEGO_KLT_ERR_STATE = -5         /* batch object used out of order (run before upload ...)      */
};

/* ---- solver kernel selection ---- */
enum {
    LEGO_KLT_KERNEL_AUTO = 0,   /* LANE where it is compiled (forward mode: patch -3..3, -4..3 or -5..5; the reference's
                                   inverse mode: -3..3) and the call has more than 3000 features, else PATCH (lowest
                                   latency on small calls) */
    LEGO_KLT_KERNEL_EXACT = 1,  /* one thread per feature, reference operation order, flat global
                                   addressing: bit-identical to the CPU oracle; the on-GPU checker  */
    LEGO_KLT_KERNEL_WARP = 2,   /* one warp per feature, windows staged in shared memory, fp64
                                   warp-shuffle reductions, all levels fused in-kernel             */
    LEGO_KLT_KERNEL_LANE = 3,   /* one thread per feature, persistent iteration state machine,
                                   shared sample grid; forward 7x7, 8x8, 11x11, inverse 7x7 */
    LEGO_KLT_KERNEL_PATCH = 4   /* one CTA per feature, one thread per patch pixel, every sample as the reference
                                   takes it: the shortest Gauss-Newton pass -- lowest latency on small calls      */
};

/*
 * Tracker parameters.  Defaults (lego_klt_default_params) are the reference's compile-time
 * literals: src/algorithm.cpp:40-42 (half_patch_size 3, half_grad_step 1, iterations 10),
 * :113 (1e-2), :135-137 (4 levels, scale 0.5); call sites pass inverse=false, has_initial=true
 * (src/frontend_g2o.cpp:473,515).
 */
typedef struct lego_klt_params {
    int32_t levels;       /* pyramid levels L (1 == LKOpticalFlow1Layer, 4 == LKOpticalFlow4Layer) */
    int32_t patch_lo;     /* patch offsets x,y run patch_lo..patch_hi inclusive; reference: -3     */
    int32_t patch_hi;     /* reference: +3  (7x7).  8x8 = (-4,3);  11x11 = (-5,5)                  */
    int32_t max_iters;    /* Gauss-Newton iterations per level; reference: 10                      */
    int32_t inverse;      /* 0 = forward-additive; 1 = the reference's inverse mode INCLUDING its
                             stale-Jacobian behaviour (src/algorithm.cpp:57,74-80,83)              */
    int32_t has_initial;  /* use kp2 as the initial guess on the coarsest level                    */
    int32_t kernel;       /* LEGO_KLT_KERNEL_*                                                     */
    int32_t reserved;
    double eps;           /* convergence: stop when |update| < eps; reference: 1e-2                */
} lego_klt_params;

/* Per-call counters (SURVEY.md 5 "Metrics", 8d: iteration counts drive the roofline accounting). */
typedef struct lego_klt_stats {
    uint64_t n_features;
    uint64_t n_success;                          /* final flags set                                */
    uint64_t n_nan;                              /* solves that returned NaN/Inf (algorithm.cpp:94) */
    uint64_t n_out_of_image;                     /* final level: result outside img2 (:123)        */
    uint64_t gn_iters[LEGO_KLT_MAX_LEVELS];      /* patch passes executed per level, level 0 = fine */
    uint64_t n_slow_path;                        /* warp kernel: passes run on the exact per-pixel path */
    uint64_t n_deferred;                         /* lane kernel: features handed to the warp kernel  */
    uint64_t defer_reason[4];                    /* [0] of n_deferred: irregular template grid (kx+c with more than
                                                    one rounding error, window outside the apron); lane kernel,
                                                    two-family levels: [1] passes on such levels, [2] of which the
                                                    family split changed a rounded coordinate (masked sub-passes),
                                                    [3] warp trips through the masked row loop              */
    float ms_h2d, ms_pyramid, ms_solver, ms_d2h; /* device times (CUDA events) of the last call    */
} lego_klt_stats;

typedef struct lego_klt_ctx lego_klt_ctx;       /* one per (calling thread, device)               */
typedef struct lego_klt_batch lego_klt_batch;   /* device-resident batch of B image pairs         */
typedef struct lego_klt_image lego_klt_image;   /* one device-resident image with its cached pyramid */

int lego_klt_abi_version(void);
/* Kernels of this library launched by this process so far (every <<<>>> is counted where it is issued). */
long long lego_klt_kernel_launches(void);
const char *lego_klt_last_error(void);           /* thread-local message of the last failure       */
void lego_klt_default_params(lego_klt_params *p);
int lego_klt_device_count(void);                 /* >=0, or LEGO_KLT_ERR_NO_DEVICE                 */

/* Context: owns a stream, scratch device buffers and pinned staging buffers on `device`.
 * Ownership: every lego_klt_batch / lego_klt_image points into the context it was created from.  lego_klt_destroy
 * while such handles are alive is allowed: the context is then only marked and is torn down when the last of its
 * batches / images is destroyed; it must not be USED (passed to any other entry point) after lego_klt_destroy. */
int lego_klt_create(int device, lego_klt_ctx **out);
void lego_klt_destroy(lego_klt_ctx *ctx);
/* Launch on a caller-owned stream (cudaStream_t passed as void*); NULL restores the context's own. */
int lego_klt_set_stream(lego_klt_ctx *ctx, void *cuda_stream);

/*
 * Replaces LKOpticalFlow4Layer / LKOpticalFlow1Layer (levels = 4 / 1).  Synchronous: returns after
 * results are in the caller's buffers.
 *   img1,img2 : 8-bit single channel, `rows` x `cols`, `step` bytes per row  (cv::Mat data/cols/rows/step)
 *   kp1_xy    : n x {x,y} float  (cv::KeyPoint::pt of kp1)
 *   kp2_xy    : n x {x,y} float, in: initial guesses (read when has_initial), out: tracked points
 *   success   : n bytes out, 0/1   (std::vector<bool> success)
 */
int lego_klt_track(lego_klt_ctx *ctx, const lego_klt_params *params,
                   const uint8_t *img1, const uint8_t *img2, int cols, int rows, size_t step,
                   const float *kp1_xy, float *kp2_xy, uint8_t *success, int n,
                   lego_klt_stats *stats_or_null);

/*
 * The pyramid part of LKOpticalFlow4Layer alone (src/algorithm.cpp:140-154): builds levels
 * 1..levels-1 of `img` on the GPU and copies them, tightly packed (step == cols) and concatenated
 * level 1 first, to `out`.  level_cols/level_rows (length `levels`) receive all level sizes.
 */
int lego_klt_build_pyramid(lego_klt_ctx *ctx, const uint8_t *img, int cols, int rows, size_t step,
                           int levels, uint8_t *out, size_t out_capacity,
                           int *level_cols, int *level_rows);

/*
 * Test hook: after lego_klt_build_pyramid, copies the DEVICE rows of `level` (0 = the uploaded image) as they
 * lie in HBM -- `rows` rows of `*pitch` bytes, pixel (r, 0) at byte r * *pitch + *apron_left -- so that the
 * row aprons the solver's border semantics rely on (algorithm.h:42-55 clamps and flat addressing) can be
 * checked directly.  Not needed by the reference's call sites.
 */
int lego_klt_debug_read_level(lego_klt_ctx *ctx, int level, uint8_t *out, size_t out_capacity,
                              int *pitch, int *apron_left);

/*
 * Sequence mode: the reference rebuilds both pyramids inside every LKOpticalFlow4Layer
 * call (src/algorithm.cpp:140-154), so in Frontend::Track the same left image is pyramided up to three
 * times (img2 of the temporal track, img1 of the stereo match, img1 of the next temporal track).  An image
 * handle uploads an image once and caches its pyramid; lego_klt_track_images runs the solver only and gives
 * the same bytes as lego_klt_track on the same two images.
 */
int lego_klt_image_create(lego_klt_ctx *ctx, int cols, int rows, size_t step, int levels, lego_klt_image **out);
void lego_klt_image_destroy(lego_klt_image *img);
/* H2D + pyramid + aprons (asynchronous on the context stream; `data` is copied before returning). */
int lego_klt_image_upload(lego_klt_image *img, const uint8_t *data);
/*
 * Image ingest: Dataset::NextFrame halves every frame before the frontend sees it,
 *     cv::resize(image, resized, cv::Size(), 0.5, 0.5, cv::INTER_NEAREST)          src/dataset.cpp:75-77
 * (OpenCV, third party: dsize = cvRound(size * 0.5), pixel (x, y) = source (min(2x, cols-1), min(2y, rows-1))).
 * Uploads the FULL-resolution frame and halves it on the device into the handle's level 0, then builds the
 * pyramid.  The handle must have been created with cols/rows = lego_klt_half_size() of the frame and step == cols
 * (cv::resize outputs are continuous).
 */
int lego_klt_image_upload_fullres(lego_klt_image *img, const uint8_t *full, int full_cols, int full_rows,
                                  size_t full_step);
/* cvRound(v * 0.5): the size cv::resize(..., Size(), 0.5, 0.5) gives an axis of v pixels. */
int lego_klt_half_size(int v);
/* The same halving, host to host (out: tightly packed half_rows x half_cols), for callers and tests. */
int lego_klt_downscale_half(lego_klt_ctx *ctx, const uint8_t *full, int full_cols, int full_rows, size_t full_step,
                            uint8_t *out, size_t out_capacity);
int lego_klt_track_images(lego_klt_ctx *ctx, const lego_klt_params *params, const lego_klt_image *img1,
                          const lego_klt_image *img2, const float *kp1_xy, float *kp2_xy, uint8_t *success, int n,
                          lego_klt_stats *stats_or_null);

/*
 * One frame of Frontend::Track in one call (SURVEY.md 8f N1 "temporal + stereo tracking of a frame as one submission"):
 * the temporal track last left -> current left (src/frontend_g2o.cpp:247-256, 453-492) and, chained on the device, the
 * stereo match current left -> current right (:299-308, 495-535) of the features the temporal track kept -- the tracked
 * positions are the stereo source points and its initial guesses (:508) without leaving HBM; one keypoint upload, one
 * synchronisation, both result groups in one read-back.  Equal, byte for byte, to lego_klt_track_images(prev, cur)
 * followed by lego_klt_track_images(cur, right) on the kept features.
 *   kp_cur_xy       : in: initial guesses of the temporal track, out: tracked positions      success_temporal : n bytes
 *   kp_right_xy     : out: matched positions in the right image (a slot whose temporal track failed: its tracked
 *                     position, success_stereo = 0, no counters)                              success_stereo   : n bytes
 * The three handles must have been uploaded (lego_klt_image_upload is asynchronous on the same stream).
 */
int lego_klt_track_frame(lego_klt_ctx *ctx, const lego_klt_params *params, const lego_klt_image *prev_left,
                         const lego_klt_image *cur_left, const lego_klt_image *cur_right, const float *kp_prev_xy,
                         float *kp_cur_xy, uint8_t *success_temporal, float *kp_right_xy, uint8_t *success_stereo, int n,
                         lego_klt_stats *stats_temporal_or_null, lego_klt_stats *stats_stereo_or_null);

/*
 * Batched path (north star (3)): B independent image pairs of one shape, n features per pair.
 * Host buffers should come from lego_klt_alloc_pinned for asynchronous copies.
 *   imgs1, imgs2 : B images, image b at  base + b * rows * step
 *   kp1_xy, kp2_xy : B*n x {x,y};  success : B*n bytes
 */
int lego_klt_batch_create(lego_klt_ctx *ctx, int batch, int cols, int rows, size_t step,
                          int n_per_pair, int levels, lego_klt_batch **out);
void lego_klt_batch_destroy(lego_klt_batch *b);
/*
 * Ragged batches: pair b tracks only its first counts[b] (0..n_per_pair) features; the remaining slots of the pair are
 * left untracked (kp2 out = kp2 in, success = 0, not counted).  counts == NULL returns to "every pair has n_per_pair
 * features".  The reference's callers pass a different number of features with every call (std::vector sizes,
 * src/frontend_g2o.cpp:457-468); a batch of such calls is ragged.  The array is copied before the call returns.
 */
int lego_klt_batch_set_feature_counts(lego_klt_batch *b, const int *counts);
/* H2D of images and keypoints (asynchronous on the context stream). */
int lego_klt_batch_upload(lego_klt_batch *b, const uint8_t *imgs1, const uint8_t *imgs2,
                          const float *kp1_xy, const float *kp2_xy);
/* Pyramids + solver on whatever is resident; inputs are preserved, so it can be re-run. */
int lego_klt_batch_run(lego_klt_batch *b, const lego_klt_params *params);
/* D2H of results; synchronises the stream. */
int lego_klt_batch_download(lego_klt_batch *b, float *kp2_xy, uint8_t *success,
                            lego_klt_stats *stats_or_null);
/* Device time of the pyramid and solver kernels, averaged over the last `last_n` (<= 64) runs of this
 * batch, from CUDA events recorded around each launch on the batch's stream.  Synchronises. */
int lego_klt_batch_timings(lego_klt_batch *b, int last_n, float *ms_pyramid_avg, float *ms_solver_avg);
/* upload + run + download in one call (the end-to-end path bench.py times as `e2e`). */
int lego_klt_track_batched(lego_klt_batch *b, const lego_klt_params *params,
                           const uint8_t *imgs1, const uint8_t *imgs2,
                           const float *kp1_xy, float *kp2_xy, uint8_t *success,
                           lego_klt_stats *stats_or_null);
/* The same in two halves (SURVEY.md 8b "batched API may be async with explicit wait"): _begin enqueues the copies and
 * kernels of the call and returns; _end waits for them (results are in the caller's buffers afterwards).  With two batch
 * objects (on two contexts) a caller keeps the copy engine busy across calls: begin(A), begin(B), end(A), begin(A) ...
 * The buffers passed to _begin must stay valid and untouched until _end. */
int lego_klt_track_batched_begin(lego_klt_batch *b, const lego_klt_params *params, const uint8_t *imgs1,
                                 const uint8_t *imgs2, const float *kp1_xy, float *kp2_xy, uint8_t *success);
int lego_klt_track_batched_end(lego_klt_batch *b, lego_klt_stats *stats_or_null);
/* lego_klt_track_batched cuts large batches into `chunks` groups of pairs whose H2D copy, kernels and D2H copy overlap
 * (1..16; 0 = the default: 6 for 32 pairs or more, 4 for 8..31, else one). */
int lego_klt_batch_set_pipeline_chunks(lego_klt_batch *b, int chunks);
/* Device pointers of the resident buffers (for callers that already hold data in HBM). */
int lego_klt_batch_device_ptrs(lego_klt_batch *b, void **imgs1, void **imgs2,
                               void **kp1_xy, void **kp2_xy_init, void **kp2_xy_out, void **success);
int lego_klt_sync(lego_klt_ctx *ctx);

void *lego_klt_alloc_pinned(size_t bytes);
void lego_klt_free_pinned(void *p);

/*
 * ---- Feature detection -------------------------------------------------------------------
 * Frontend::DetectFeatures (src/frontend_g2o.cpp:279-297): cv::GFTTDetector::create(num_features, 0.01, 20) (:16) on the
 * left image, under a mask that is 0 in the rectangle pt +- (10, 10) around every feature the frame already has
 * (:280-284).  OpenCV's goodFeaturesToTrack (third party: Shi-Tomasi minimum eigenvalue of the 3x3 structure tensor of
 * Sobel derivatives, threshold at quality_level * max, 3x3 non-maximum suppression, greedy minimum-distance selection in
 * score order) restated on the device; pinned against Python cv2 to a stated tolerance (tests/test_gftt.py).
 *   mask (may be NULL)        rows x cols bytes, mask_step per row: candidates only where != 0      (cv::Mat mask)
 *   exclude_xy (may be NULL)  n_exclude x {x, y}: the mask is additionally cleared in pt +- exclude_half around each, both
 *                             corners inclusive (cv::rectangle(mask, pt - (h,h), pt + (h,h), 0, CV_FILLED))
 *   corners_xy                max_corners x {x, y} out, in OpenCV's order (by decreasing score); scores (may be NULL)
 *   *n_corners                number of corners found (<= max_corners)
 */
int lego_klt_detect_features(lego_klt_ctx *ctx, const uint8_t *img, int cols, int rows, size_t step, const uint8_t *mask,
                             size_t mask_step, const float *exclude_xy, int n_exclude, float exclude_half, int max_corners,
                             double quality_level, double min_distance, float *corners_xy, float *scores_or_null,
                             int *n_corners);
/* The same on an uploaded image handle (its level 0 where it lies in HBM): only the exclusion list goes up and the
 * corners come back. */
int lego_klt_image_detect_features(lego_klt_image *img, const float *exclude_xy, int n_exclude, float exclude_half,
                                   int max_corners, double quality_level, double min_distance, float *corners_xy,
                                   float *scores_or_null, int *n_corners);
/* The same for every pair of a batch whose images are in HBM (lego_klt_batch_upload / lego_klt_track_batched): one
 * Frontend::DetectFeatures per image of `set` (0 = img1, 1 = img2), all pairs per launch.  exclude_source_keypoints != 0:
 * the mask of src/frontend_g2o.cpp:280-284 is built from the pair's source keypoints (the first counts[b] of them in a
 * ragged batch).  corners_xy: B x max_corners x {x, y} in OpenCV's order (slots beyond n_corners[b] are zero);
 * scores_or_null: B x max_corners; n_corners: B. */
int lego_klt_batch_detect_features(lego_klt_batch *b, int set, int exclude_source_keypoints, float exclude_half,
                                   int max_corners, double quality_level, double min_distance, float *corners_xy,
                                   float *scores_or_null, int *n_corners);
/* Makes the corners of the batch's last lego_klt_batch_detect_features call its source keypoints, where they lie in HBM:
 * pair b tracks its n_corners[b] corners (a ragged batch), initial guess = the same pixel (src/frontend_g2o.cpp:508).
 * Then lego_klt_batch_run / lego_klt_batch_triangulate: detection -> stereo matching -> triangulation of a batch of
 * frames without the keypoints crossing PCIe.  The detection's max_corners must not exceed the batch's n_per_pair.
 * The per-pair counts stay in force like those of lego_klt_batch_set_feature_counts (reset them with counts = NULL). */
int lego_klt_batch_use_detected_features(lego_klt_batch *b);
/* Test hook: the minimum-eigenvalue map (rows x cols floats) of the context's last detection. */
int lego_klt_debug_read_eig(lego_klt_ctx *ctx, float *out, size_t capacity, int *cols, int *rows);

/*
 * One process, several devices (SURVEY.md 8b "device list", 8e): the B pairs are cut into contiguous blocks, one per
 * entry of `devices` (block sizes differ by at most one; a device may be listed more than once), each block owned by
 * a context and a batch on its device.  lego_klt_multi_track runs lego_klt_track_batched on every block concurrently,
 * one host thread per device, on slices of the caller's (pinned) buffers; pairs are independent, so there is no
 * exchange between devices and the result bytes equal those of a single-device call on the whole batch.
 */
typedef struct lego_klt_multi lego_klt_multi;
int lego_klt_multi_create(const int *devices, int n_devices, int batch, int cols, int rows, size_t step,
                          int n_per_pair, int levels, lego_klt_multi **out);
void lego_klt_multi_destroy(lego_klt_multi *m);
/* Block i: its device, first pair and pair count.  Returns the number of blocks (or a negative error). */
int lego_klt_multi_shard(const lego_klt_multi *m, int i, int *device, int *first_pair, int *n_pairs);
int lego_klt_multi_set_feature_counts(lego_klt_multi *m, const int *counts);   /* B entries, or NULL */
/*
 * Schedule of lego_klt_multi_track.  block_pairs = 0 (default): the static contiguous blocks above.  block_pairs > 0:
 * the devices PULL blocks of that many pairs from one shared counter, two blocks in flight per device (upload of one
 * behind the kernels and the result copy of the other), so that a device behind a slower host link -- or one that is
 * shared with other work -- takes fewer pairs instead of holding the call back.  Same result bytes either way.
 * lego_klt_multi_last_distribution: pairs each device tracked in the last call (returns the number of devices).
 */
int lego_klt_multi_set_schedule(lego_klt_multi *m, int block_pairs);
int lego_klt_multi_last_distribution(const lego_klt_multi *m, int *pairs_per_device, int capacity);
int lego_klt_multi_track(lego_klt_multi *m, const lego_klt_params *params, const uint8_t *imgs1,
                         const uint8_t *imgs2, const float *kp1_xy, float *kp2_xy, uint8_t *success,
                         lego_klt_stats *stats_or_null);

/*
 * ---- Triangulation of tracked features ---------------------------------------------------
 * legoslam::triangulation (include/legoslam/algorithm.h:11-34), called per feature right after
 * FindFeaturesInRight by Frontend::TriangulateNewPoints / BuildInitMap (src/frontend_g2o.cpp:111-155, :310-349)
 * with poses = {camera_left->pose(), camera_right->pose()}.  One GPU thread per feature; same outputs: the world
 * point (written whatever the verdict, like the reference's out-parameter) and the bool it returns
 * (finite && S[3]/S[2] < sing_ratio_thr).  The caller's extra gates (y <= 2 m, depth limits) stay with the caller.
 */
#define LEGO_TRI_MAX_VIEWS 8

/* Camera (include/legoslam/camera.h:13-24): intrinsics and pose_.matrix3x4(), row-major. */
typedef struct lego_camera {
    double fx, fy, cx, cy;
    double pose34[12];
} lego_camera;

/*
 * Generic form: n features seen in the same n_views (2..LEGO_TRI_MAX_VIEWS) views.
 *   poses34   : n_views x 12, SE3::matrix3x4() row-major                  (std::vector<SE3> poses)
 *   points_xy : n x n_views x {x, y}: points[i][0], points[i][1]          (VecVec3 points; [2] is not read)
 *   pt_world  : n x 3 out;  ok : n bytes out (the reference's return value)
 */
int lego_klt_triangulate(lego_klt_ctx *ctx, const double *poses34, int n_views, const double *points_xy, int n,
                     double sing_ratio_thr, double *pt_world, uint8_t *ok);

/*
 * Stereo form fused with Camera::pixel2camera (src/camera.cpp:21-25, depth 1): pixel keypoints as
 * lego_klt_track* leaves them.  valid (may be null) = the tracker's success flags: a feature with valid == 0
 * has no right feature in the reference (src/frontend_g2o.cpp:114-115) and gets ok = 0, point = 0.
 */
int lego_klt_triangulate_stereo(lego_klt_ctx *ctx, const lego_camera *left, const lego_camera *right,
                            const float *kp_left_xy, const float *kp_right_xy, const uint8_t *valid, int n,
                            double sing_ratio_thr, double *pt_world, uint8_t *ok);

/*
 * The same on a batch that has been run (lego_klt_batch_run / lego_klt_track_batched): left keypoints, tracked
 * right keypoints and success flags are taken where they lie in HBM; only the results cross PCIe.
 *   pt_world : B*n x 3 out;  ok : B*n bytes out
 */
int lego_klt_batch_triangulate(lego_klt_batch *b, const lego_camera *left, const lego_camera *right,
                               double sing_ratio_thr, double *pt_world, uint8_t *ok);

#ifdef __cplusplus
}
#endif
#endif /* LEGO_KLT_H */
