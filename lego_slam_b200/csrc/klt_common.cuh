// klt_common.cuh -- shared host/device declarations of the sm_100a KLT path.
//
// Reference text restated by the device functions here (paths relative to the upstream tree):
//   include/legoslam/algorithm.h:40-57  GetPixelValue      -> sample_flat()
//   include/legoslam/algorithm.h:60-66  IsPtInImg          -> point_in_image()
//   src/algorithm.cpp:93                H.ldlt().solve(b)  -> ldlt2_solve()  (Eigen 3.3 LDLT.h, 2x2)
//
// Arithmetic contract: fp32 sampling with individually rounded mul/add (this TU is
// compiled with -fmad=false; fused ops are written explicitly as fma() only where the product is
// exact), fp64 for dx,dy,J,H,b,cost,update.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/lego_klt.h"

namespace legoklt {

// Kernel launches issued by this library since it was loaded (lego_klt_kernel_launches): every launch_* wrapper
// calls note_launch() once per <<<>>> it issues.  Defined in lego_klt_capi.cu.
extern std::atomic<long long> g_kernel_launches;
inline void note_launch(int n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

constexpr int kMaxLevels = LEGO_KLT_MAX_LEVELS;
constexpr int kMaxPatch = 13;       // (patch_hi - patch_lo + 1) <= 13
constexpr int kWinW = 48;           // TMA window box, bytes per row (multiple of 16)
constexpr int kWinH = 24;           // TMA window box, rows
constexpr int kApronL = 32;         // bytes of left apron in front of every device image row
constexpr int kApronR = 48;         // minimum bytes of right apron (column `cols` onwards)

// One pyramid level of a batch of images, device resident.
//   pixel (img k, row r, col c) of set s lives at base[s] + k*slot + r*pitch + c
// Every row carries an apron (written by the apron kernel after the pyramid build), valid for
// c in [-kApronL, pitch - kApronL):  c < 0 replicates column 0;  c == cols holds the byte the reference's
// flat addressing reads one past the row end, data[r*step + cols] (algorithm.h:48,53: first pixel of the
// next row, 0 past the last row);  c > cols replicates column cols-1.
// `step` is the LOGICAL row step the reference would see (level 0: the caller's cv::Mat::step,
// levels >= 1: cols, because cv::resize outputs are continuous) -- the border path reproduces the
// reference's flat addressing data[int(y)*step + int(x) (+1, +step, +step+1)] with it.
struct LevelView {
    uint8_t *base[2];
    unsigned long long slot;  // bytes between consecutive images (multiple of 256)
    int cols, rows;
    int pitch;                // device bytes per row incl. aprons (multiple of 16)
    int step;                 // logical step (see above)
};

struct PyramidView {
    LevelView lv[kMaxLevels];
    int levels;
    int n_images;             // images per set (= batch size B)
};

// Device-side counters (one block of uint64 per batch).
enum StatSlot {
    kStatIters0 = 0,                       // + level
    kStatNan = kMaxLevels,
    kStatOutOfImage,
    kStatSuccess,
    kStatSlowPath,
    kStatTmaTimeout,
    kStatDeferred,
    kStatDeferInexact,   // features deferred because some level's template grid is irregular
    kStatFamPasses,      // lane kernel: passes on a level with two coordinate families on some axis
    kStatFamSplit,       // ... of which the split changed a rounded coordinate (extra, masked trips)
    kStatMaskedTrips,    // lane kernel: warp trips through the masked row loop
    kStatCount
};

struct SolverArgs {
    const float2 *kp1;        // [B*n]
    const float2 *kp2_init;   // [B*n]
    float2 *kp2_out;          // [B*n]
    uint8_t *success;         // [B*n]
    unsigned long long *stats;  // [kStatCount]
    int n_per_pair;           // feature slots per pair (the batch's stride); image = id / n_per_pair
    const int *pair_count;    // [B] valid features of each pair (<= n_per_pair), or null = all: a slot at or beyond its
                              // pair's count is not tracked (kp2_out = kp2_init, success = 0, no counters)
    const uint8_t *slot_valid;  // [B*n] or null: a slot whose byte is 0 is not tracked either (same outputs) -- the fused
                              // frame call chains the stereo match on the features the temporal track kept
    int n_total;              // features in this launch: global ids f0 .. f0 + n_total - 1
    int f0;                   // first global feature id (chunked batches); image = id / n_per_pair
    int patch_lo, patch_hi;
    int max_iters;
    int inverse;
    int has_initial;
    double eps;
    double eps_sq;            // smallest double whose correctly rounded sqrt is >= eps: sqrt(x) < eps  <=>  x < eps_sq
                              // (sqrt is monotone), so the fast kernels test |update|^2 and skip the fp64 sqrt
    float one;                // 1.0f, opaque to the compiler (see add2_product in klt_solver_lane.cu)
    int debug_flags;          // LEGO_KLT_DEBUG env: 1 = never use the TMA fast path, 2 = count TMA timeouts instead of trapping
    // Optional work list (used for the features the LANE kernel defers): feature ids and their count.
    const int *list;
    const int *list_count;
    // LANE kernel work distribution / deferral (device scalars, zeroed before each run).
    double *scratch;          // LANE kernel: 6 doubles per resident thread (multi-family partial sums)
    float *templates;         // [levels][tpl_features][52]: I1 patches + regularity flag (template kernel), level-major
    unsigned long long tpl_features;  // features per level in `templates` (the batch's capacity)
    int *feat_flag;           // [n_total]: 4*epoch+2 -> warp kernel owns the feature, 4*epoch+1 -> lane<FAMILIES>
    int epoch;                // run counter (>= 1): flags from earlier runs are stale, no memset needed
    int *work_counter;
    int *defer_list;
    int *defer_count;
    int *fam_list;            // features with two coordinate families on some level (template kernel -> lane<true>)
    int *fam_count;
};

#ifdef __CUDACC__

// ------------------------------------------------------------------------------------------------
// Flat-addressed byte fetch with the reference's semantics on our pitched layout.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fetch_flat(const uint8_t *img, const LevelView &lv, long long f) {
    int row = (int)(f / lv.step);
    int col = (int)(f - (long long)row * lv.step);
    if (row >= lv.rows) return 0.f;        // reference: out-of-buffer read; oracle: zero padding
    return (float)__ldg(img + (size_t)row * lv.pitch + col);
}

// GetPixelValue -- algorithm.h:40-57.  Products and sums individually rounded, source order.
__device__ __forceinline__ float sample_flat(const uint8_t *img, const LevelView &lv, float x, float y) {
    if (x < 0.f) x = 0.f;
    if (y < 0.f) y = 0.f;
    if (x >= (float)lv.cols) x = (float)(lv.cols - 1);
    if (y >= (float)lv.rows) y = (float)(lv.rows - 1);
    int ix = (int)x, iy = (int)y;
    float xx = x - floorf(x);
    float yy = y - floorf(y);
    float p0, p1, p2, p3;
    if (ix + 1 < lv.cols && iy + 1 < lv.rows) {   // all four taps inside the image: plain 2-D fetch
        const uint8_t *p = img + (size_t)iy * lv.pitch + ix;
        p0 = (float)__ldg(p);
        p1 = (float)__ldg(p + 1);
        p2 = (float)__ldg(p + lv.pitch);
        p3 = (float)__ldg(p + lv.pitch + 1);
    } else {                                      // sliver / last row: the reference's flat addressing
        long long f = (long long)iy * lv.step + ix;
        p0 = fetch_flat(img, lv, f);
        p1 = fetch_flat(img, lv, f + 1);
        p2 = fetch_flat(img, lv, f + lv.step);
        p3 = fetch_flat(img, lv, f + lv.step + 1);
    }
    float omx = 1.f - xx, omy = 1.f - yy;
    float r = __fmul_rn(__fmul_rn(omx, omy), p0);
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xx, omy), p1));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(omx, yy), p2));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xx, yy), p3));
    return r;
}

// IsPtInImg -- algorithm.h:60-66.
__device__ __forceinline__ bool point_in_image(float px, float py, const LevelView &lv) {
    double x = px, y = py;
    return !(x < 0 || y < 0 || x >= lv.cols || y >= lv.rows);
}

// Eigen 3.3 LDLT<Matrix2d,Lower> compute()+solve() for 2x2 (third party, restated):
// pivot on the larger |diagonal| (first on ties); all-zero diagonal stops the factorisation;
// solve uses the pseudo-inverse of D with tol = 1/DBL_MAX.  No fused multiply-adds (the CPU
// reference build has none).
__device__ __forceinline__ void ldlt2_solve(double h00, double h10, double h11, double b0, double b1,
                                            double &x0, double &x1) {
    bool swapped = fabs(h11) > fabs(h00);
    double a = swapped ? h11 : h00;
    double d = swapped ? h00 : h11;
    double l, d1;
    if (a == 0.0) {
        l = h10;
        d1 = d;
    } else {
        l = __ddiv_rn(h10, a);
        d1 = __dsub_rn(d, __dmul_rn(l, __dmul_rn(a, l)));
    }
    double y0 = swapped ? b1 : b0;
    double y1 = swapped ? b0 : b1;
    y1 = __dsub_rn(y1, __dmul_rn(l, y0));
    const double tol = 1.0 / 1.7976931348623157e308;
    y0 = (fabs(a) > tol) ? __ddiv_rn(y0, a) : 0.0;
    y1 = (fabs(d1) > tol) ? __ddiv_rn(y1, d1) : 0.0;
    y0 = __dsub_rn(y0, __dmul_rn(l, y1));
    x0 = swapped ? y1 : y0;
    x1 = swapped ? y0 : y1;
}

__device__ __forceinline__ bool not_finite(double v) { return isnan(v) || isinf(v); }

// Ragged batches: is feature slot f (of image img) beyond its pair's feature count?
__device__ __forceinline__ bool slot_unused(const SolverArgs &args, int f, int img) {
    if (args.slot_valid != nullptr && args.slot_valid[f] == 0) return true;   // (written by an earlier kernel: no __ldg)
    return args.pair_count != nullptr && (f - img * args.n_per_pair) >= __ldg(args.pair_count + img);
}

// What an unused slot receives (once, by whichever kernel meets it first).
__device__ __forceinline__ void write_unused_slot(const SolverArgs &args, int f) {
    args.kp2_out[f] = args.kp2_init[f];
    args.success[f] = 0;
}

#endif  // __CUDACC__

}  // namespace legoklt
