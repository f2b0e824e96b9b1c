// klt_kernels.h -- host-side launch interface of the sm_100a kernels (internal to the library).
#pragma once

#include "klt_common.cuh"

namespace legoklt {

// ---- pyramid (pyramid_sm100.cu) ---------------------------------------------------------------
// Fixed-point resize tables of one level transition (l-1 -> l), restating OpenCV's
// cv::resize INTER_LINEAR for CV_8UC1 (third party; reference call sites src/algorithm.cpp:147-150).
struct ResizeTables {
    int *xofs;      // [cols_l]  left tap in level l-1
    short2 *xcoef;  // [cols_l]  {a0, a1}, 2^11 scale
    int *yofs;      // [rows_l]  top tap
    short2 *ycoef;  // [rows_l]
    int *gofs;      // [ceil(cols_l/4)]  first tap of a regular 4-pixel group (taps sx0 + 2j), else -1
    int *cofs;      // [ceil(cols_l/8)]  same for the 8-pixel chunks of the level 0 -> 1 streaming kernel
    int *girr;      // the groups with gofs < 0, and their number
    int n_girr;
    int *irr;       // the chunks with cofs < 0, and their number
    int n_irr;
    int x_exact2;   // xofs[c] == 2c and a0 == a1 == 1024 for every c
    int y_exact2;   // same for the rows
};

struct PyramidPlan {
    int levels = 0;
    int cols[kMaxLevels] = {0}, rows[kMaxLevels] = {0}, pitch[kMaxLevels] = {0};
    ResizeTables tab[kMaxLevels] = {};  // tab[l] valid for l >= 1 (device pointers)
    int top_rows_per_cta = 1;
    int smem_off[kMaxLevels] = {0};     // byte offset of level l's staging rows in dynamic smem
    int max_rows[kMaxLevels] = {0};     // max staged rows of level l over all CTAs
    size_t smem_bytes = 0;
    void *table_blob = nullptr;         // single device allocation behind all tables
    void *band_tab = nullptr;           // device int2 [n_bands][kMaxLevels]: {first row, rows} of every level's band
    int n_bands = 0;
    // Level 0 -> 1 runs as a streaming kernel; levels 2.. are built by the band kernel from level 1, described by
    // a plan of their own (the same pyramid seen from level 1).  Null in such a sub-plan and when levels == 1.
    PyramidPlan *sub = nullptr;
};

// Level sizes of the reference's pyramid: cv::Size(cols*0.5, rows*0.5) repeatedly (truncation).
// Returns false if a level would be empty.
bool pyramid_level_sizes(int cols, int rows, int levels, int *lcols, int *lrows);
cudaError_t pyramid_plan_create(int cols, int rows, int levels, const int *pitch, PyramidPlan *plan,
                                bool band_kernel_only = false);
void pyramid_plan_destroy(PyramidPlan *plan);
// Builds levels 1..L-1 of images [img0, img0 + nimg) of both image sets from level 0 AND writes the row aprons
// (see LevelView) of every level, one fused launch.
cudaError_t launch_pyramid(const PyramidPlan &plan, const PyramidView &pyr, int img0, int nimg, cudaStream_t stream,
                           int n_sets = 2);
// Re-pitches tight level-0 rows (`step` bytes per row, all images of the set back to back starting at the
// 4-byte aligned `tight`, buffer padded by >= 8 bytes) of images [img0, img0+n_images) into the device layout.
cudaError_t launch_ingest(const uint8_t *tight, const LevelView &l0, int set, int img0, int n_images,
                          cudaStream_t stream);
// 0.5x INTER_NEAREST downscale of a tight full-resolution frame (device) into level 0 of a single device image
// (Dataset::NextFrame, src/dataset.cpp:75-77).
cudaError_t launch_half_nearest(const uint8_t *full, int full_cols, int full_rows, size_t full_step, const LevelView &l0,
                                cudaStream_t stream);
// Row aprons alone (single-level configurations; launch_pyramid calls it when there is no level to build).
cudaError_t launch_aprons(const PyramidView &pyr, int img0, int nimg, cudaStream_t stream, int n_sets = 2);

// ---- solver kernels ---------------------------------------------------------------------------
cudaError_t launch_klt_exact(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream);
// One CTA per feature, one thread per patch pixel (klt_solver_patch.cu): the low-latency kernel of small calls.
bool patch_kernel_supports(const SolverArgs &args);
cudaError_t launch_klt_patch(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream);

struct WarpKernelMaps;  // TMA descriptors, defined in klt_solver_warp.cu
cudaError_t warp_maps_create(const PyramidView &pyr, WarpKernelMaps **out);
void warp_maps_destroy(WarpKernelMaps *maps);
cudaError_t launch_klt_warp(const PyramidView &pyr, const WarpKernelMaps *maps, const SolverArgs &args,
                            int sm_count, cudaStream_t stream);

// LANE path (klt_solver_lane.cu), 7x7 forward only: launch_klt_template computes the I1 patches and puts
// irregular features on args.defer_list (solve them with launch_klt_warp, args.list = defer_list, which may
// run concurrently with launch_klt_lane: the two touch disjoint features).
bool lane_kernel_supports(const SolverArgs &args);
size_t lane_template_bytes(int n_total, int levels);
size_t lane_scratch_bytes(int sm_count);
cudaError_t launch_klt_template(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream);
cudaError_t launch_klt_lane(const PyramidView &pyr, const SolverArgs &args, int sm_count, cudaStream_t stream,
                            cudaStream_t families_stream);
// The same path compiled for the 8x8 patch, offsets -4..3 (klt_solver_lane_p8.cu).
bool lane_kernel_supports_p8(const SolverArgs &args);
size_t lane_template_bytes_p8(int n_total, int levels);
size_t lane_scratch_bytes_p8(int sm_count);
cudaError_t launch_klt_template_p8(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream);
cudaError_t launch_klt_lane_p8(const PyramidView &pyr, const SolverArgs &args, int sm_count, cudaStream_t stream,
                            cudaStream_t families_stream);
// ... and for the 11x11 patch, offsets -5..5 (klt_solver_lane_p11.cu).
bool lane_kernel_supports_p11(const SolverArgs &args);
size_t lane_template_bytes_p11(int n_total, int levels);
size_t lane_scratch_bytes_p11(int sm_count);
cudaError_t launch_klt_template_p11(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream);
cudaError_t launch_klt_lane_p11(const PyramidView &pyr, const SolverArgs &args, int sm_count, cudaStream_t stream,
                            cudaStream_t families_stream);

// ... and for the reference's inverse mode, 7x7 patch (klt_solver_lane_inv.cu; BASELINE config C4).
bool lane_kernel_supports_inv(const SolverArgs &args);
size_t lane_template_bytes_inv(int n_total, int levels);
size_t lane_scratch_bytes_inv(int sm_count);
cudaError_t launch_klt_template_inv(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream);
cudaError_t launch_klt_lane_inv(const PyramidView &pyr, const SolverArgs &args, int sm_count, cudaStream_t stream,
                                cudaStream_t families_stream);

// ---- triangulation (triangulate_sm100.cu) ------------------------------------------------
constexpr int kTriMaxViews = 8;
// poses34: host, n_views x 12 (row-major 3x4); points: device, n x n_views x {x, y}; outputs: device n x 3 / n.
cudaError_t launch_triangulate(const double *poses34, int n_views, const double *d_points, int n, double thr,
                               double *d_pt_world, uint8_t *d_ok, cudaStream_t stream);
// cam_*: host {fx, fy, cx, cy}; poses34: host 2 x 12 (left, right); keypoints: device packed float2 (pixels).
cudaError_t launch_triangulate_stereo(const double *poses34, const double *cam_left, const double *cam_right,
                                      const float2 *d_kp_left, const float2 *d_kp_right, const uint8_t *d_valid, int n,
                                      double thr, double *d_pt_world, uint8_t *d_ok, cudaStream_t stream);

// ---- feature detection (gftt_sm100.cu) -----------------------------------------------------
// cv::goodFeaturesToTrack (Shi-Tomasi, blockSize 3) on a device image of `pitch` bytes per row.  d_mask_in (tight,
// cols bytes per row) and / or an exclusion list (0 in pt +- half around each point) restrict the candidates.  `ws` is a
// device workspace of gftt_workspace_bytes(); corners / scores / n_out are device outputs (corners in acceptance order).
// Synchronises the stream once (the candidate count is needed on the host to size the sort).
size_t gftt_workspace_bytes(int cols, int rows, float min_distance);
cudaError_t launch_gftt(const uint8_t *d_img, int cols, int rows, int pitch, const uint8_t *d_mask_in, const float2 *d_exclude,
                        int n_exclude, float exclude_half, int max_corners, double quality, float min_distance, uint8_t *ws,
                        size_t ws_bytes, float2 *d_corners, float *d_scores_or_null, int *d_n_out, int *h_n_candidates,
                        cudaStream_t stream);
// Batched form: n_images images of one layout (image i at d_imgs + i * img_stride, rows of `pitch` bytes), `chunk` of them
// per pass through the workspace of gftt_batched_workspace_bytes(); an exclusion list of exclude_per_image points per image
// (d_exclude_counts: how many of them are used, or null = all).  Outputs: corners [n_images][max_corners], n_out [n_images].
size_t gftt_batched_workspace_bytes(int cols, int rows, int chunk);
bool gftt_batched_supported(int cols, int rows, int batch);
cudaError_t launch_gftt_batched(const uint8_t *d_imgs, size_t img_stride, int pitch, int cols, int rows, int n_images,
                                const float2 *d_exclude, int exclude_per_image, const int *d_exclude_counts, float exclude_half,
                                int max_corners, double quality, float min_distance, uint8_t *ws, size_t ws_bytes, int chunk,
                                float2 *d_corners, float *d_scores_or_null, int *d_n_out, cudaStream_t stream);
// Test hook: the eigenvalue map of the last launch_gftt on this workspace (rows x cols floats) to host memory.
cudaError_t gftt_debug_eig(const uint8_t *ws, int cols, int rows, float *h_eig, cudaStream_t stream);

}  // namespace legoklt
