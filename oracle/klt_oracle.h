/*
 * klt_oracle.h -- CPU oracle for the pyramid Gauss-Newton KLT path of LEGO-SLAM.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lego_slam_b200/, include/) may include,
 * link or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * It restates, operation for operation, the reference's
 *     include/legoslam/algorithm.h:40-66   GetPixelValue, IsPtInImg
 *     src/algorithm.cpp:11-31              LKOpticalFlow1Layer  (parallel-for over features)
 *     src/algorithm.cpp:37-125             LKOpticalFlowTracker::calcLKOpticalFlow
 *     src/algorithm.cpp:128-206            LKOpticalFlow4Layer  (pyramids + coarse-to-fine driver)
 * plus two third-party algorithms whose source is NOT under /root/reference:
 *     OpenCV (>=3.2, un-pinned, CMakeLists.txt:18) cv::resize INTER_LINEAR on CV_8UC1
 *     Eigen3 (un-pinned, CMakeLists.txt:15; 3.3.x on the reference's Ubuntu 18.04)
 *            Matrix2d::ldlt().solve()
 *
 * PARITY PINNING: the reference ships no test, golden vector or fixture for this path, so the
 * pin is the reference ITSELF run here: oracle/_ref/libklt_ref.so is the reference's own translation unit
 * (/root/reference/src/algorithm.cpp + include/legoslam/algorithm.h, compiled unmodified where they lie, with the
 * reference's flags, on stand-in OpenCV/Eigen headers -- oracle/build_ref.py, oracle/ref_stubs/), and
 * tests/test_oracle_vs_ref.py asserts this restatement equals it BITWISE (positions, flags) on the BASELINE config
 * shapes, both modes, has_initial on/off, 1 and 4 layers, padded steps, border / sliver / outside / sub-pixel
 * points, flat and white-noise images; the committed golden vectors (tests/golden/) are _ref's outputs.
 * The resize (OpenCV, third party) is pinned against Python cv2.resize (bit exact, tests/test_oracle_pyramid.py).
 * What remains a restatement of third-party code with no pin available in this image: Eigen 3.3's 2x2 pivoted LDLT
 * (written twice, independently structured: ldlt2_solve here, LDLT<N> in ref_stubs/eigen_stub.hpp).
 *
 * Deliberate deviations from the reference (none changes results on in-range data):
 *   - success flags are bytes, not std::vector<bool> (the reference races on the bit-packed vector)
 *   - images live in buffers with >= step+2 zero bytes after the last row, so the reference's
 *     out-of-buffer reads for y in (rows-1, rows) (algorithm.h:48-55) are defined: they read 0
 *   - no stdout print on NaN; levels / patch bounds / iteration cap / eps are parameters whose
 *     defaults are the reference's literals; iteration counters are exported.
 */
#ifndef KLT_ORACLE_H
#define KLT_ORACLE_H

#include "../include/lego_klt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* cv::resize(src, dst, Size(int(sw*0.5), int(sh*0.5))) for CV_8UC1, INTER_LINEAR; dst is tight. */
int klt_oracle_resize_half(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst);

/* levels 1..levels-1, tight, concatenated (same contract as lego_klt_build_pyramid). */
int klt_oracle_build_pyramid(const uint8_t *img, int cols, int rows, size_t step, int levels,
                             uint8_t *out, size_t out_capacity, int *level_cols, int *level_rows);

/* GetPixelValue on a caller buffer of buf_len bytes; bytes at index >= buf_len read as 0. */
float klt_oracle_get_pixel_value(const uint8_t *data, int cols, int rows, size_t step,
                                 size_t buf_len, float x, float y);

/* Eigen::Matrix2d::ldlt().solve(): H row-major {h00,h01,h10,h11} (lower triangle used). */
void klt_oracle_ldlt2_solve(const double H[4], const double b[2], double x[2]);

/*
 * LKOpticalFlow4Layer (params->levels == 4) / LKOpticalFlow1Layer (levels == 1) on the CPU.
 * `threads` contiguous feature stripes per level (cv::parallel_for_ shape); <=1 runs serially.
 * params->kernel is ignored.  Returns 0 or LEGO_KLT_ERR_*.
 */
int klt_oracle_track(const lego_klt_params *params, const uint8_t *img1, const uint8_t *img2,
                     int cols, int rows, size_t step, const float *kp1_xy, float *kp2_xy,
                     uint8_t *success, int n, int threads, lego_klt_stats *stats_or_null);

/* Solver only, on prebuilt pyramids (for the solver-only CPU timing): pyr1/pyr2 are arrays of
 * `levels` pointers to tight level images whose buffers carry >= cols+2 bytes of zero padding. */
int klt_oracle_track_prebuilt(const lego_klt_params *params, const uint8_t *const *pyr1,
                              const uint8_t *const *pyr2, const int *level_cols,
                              const int *level_rows, const size_t *level_step,
                              const float *kp1_xy, float *kp2_xy, uint8_t *success, int n,
                              int threads, lego_klt_stats *stats_or_null);

#ifdef __cplusplus
}
#endif
#endif
