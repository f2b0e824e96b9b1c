// klt_solver_patch.cu -- LEGO_KLT_KERNEL_PATCH: one CTA per feature, one thread per patch pixel.
//
// The kernel for SMALL calls -- the reference's own call pattern, one pair with 150..2000 features
// (Frontend::TrackLastFrame / FindFeaturesInRight, src/frontend_g2o.cpp:473,515): there the time of a call is the
// length of the serial Gauss-Newton chain of its slowest feature (4 levels x up to 10 passes), so what counts is
// the latency of ONE pass.  Every thread evaluates its pixel exactly as the reference does (src/algorithm.cpp:63-88
// through sample_flat = GetPixelValue, algorithm.h:40-57: template sample, error, the four gradient taps), the six
// sums of a pass are reduced over the CTA (warp shuffles, then the warps in order), thread 0 does the 2x2 solve and
// the convergence logic (:92-117) and publishes the new displacement.  A pass is ~500 dependent instructions instead
// of the warp kernel's ~850 (which spreads the 81 shared samples over 32 lanes in three rounds).
//
// Fidelity: every fp32 / fp64 value entering the sums is the reference's, bit for bit -- there is no shared sample
// grid here, hence no condition to check and no feature to defer; only the ORDER of the fp64 additions differs (a
// tree instead of x-outer / y-inner), the same contract as the warp and lane kernels.  Both modes (the inverse mode
// with its stale Jacobian, :57,74-80), any patch up to 16 x 16, any level count.
#include "klt_common.cuh"
#include "klt_kernels.h"

namespace legoklt {

namespace {

constexpr int kPatchMaxThreads = 256;

struct PatchCtl {
    float kx, ky;     // source keypoint at this level
    double dx, dy;    // displacement the next pass samples at
    double j0, j1;    // inverse mode: the stale Jacobian (last pixel of the first pass)
    int cmd;          // 1: run a pass, 0: level done
    int level;        // level the next pass runs on (-1: feature done)
};

__global__ void __launch_bounds__(kPatchMaxThreads)
klt_patch_kernel(const __grid_constant__ PyramidView pyr, const __grid_constant__ SolverArgs args) {
    __shared__ double part[kPatchMaxThreads / 32][6];
    __shared__ PatchCtl ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = (blockDim.x + 31) >> 5;
    const int i = args.f0 + blockIdx.x;
    const int img = i / args.n_per_pair;
    if (slot_unused(args, i, img)) {   // (uniform over the CTA)
        if (tid == 0) write_unused_slot(args, i);
        return;
    }
    const int L = pyr.levels;
    const bool inverse = args.inverse != 0;
    const int P = args.patch_hi - args.patch_lo + 1;
    const bool active = tid < P * P;
    const int px = args.patch_lo + tid / P, py = args.patch_lo + tid % P;   // x outer, y inner like the reference
    const bool last_pixel = tid == P * P - 1;

    // thread 0: the driver's state (src/algorithm.cpp:158-205)
    float2 k1 = make_float2(0.f, 0.f), k2 = k1;
    bool flag = true;
    if (tid == 0) {
        const double scale_top = 1.0 / (double)(1 << (L - 1));   // :160-169  Point2f *= double
        k1 = args.kp1[i];
        k2 = args.kp2_init[i];
        k1.x = (float)(k1.x * scale_top);
        k1.y = (float)(k1.y * scale_top);
        k2.x = (float)(k2.x * scale_top);
        k2.y = (float)(k2.y * scale_top);
        const bool has_initial = args.has_initial != 0;           // :185-189 (top level only)
        ctl.kx = k1.x;
        ctl.ky = k1.y;
        ctl.dx = has_initial ? (double)(k2.x - k1.x) : 0.0;       // :47-50
        ctl.dy = has_initial ? (double)(k2.y - k1.y) : 0.0;
        ctl.cmd = 1;
        ctl.level = L - 1;
    }
    __syncthreads();

    int level = L - 1;
    while (level >= 0) {
        const LevelView &lv = pyr.lv[level];
        const uint8_t *img1 = lv.base[0] + (size_t)img * lv.slot;
        const uint8_t *img2 = lv.base[1] + (size_t)img * lv.slot;
        const float kx = ctl.kx, ky = ctl.ky;
        const float fx = kx + (float)px, fy = ky + (float)py;     // float adds (:65-66)
        const float i1 = active ? sample_flat(img1, lv, fx, fy) : 0.f;
        double J0 = 0, J1 = 0;
        if (inverse && active) {   // :74-80, first pass of the level: the gradient of img1 at the template position
            J0 = -1.0 * (0.5 * (double)(sample_flat(img1, lv, fx + 1.f, fy) - sample_flat(img1, lv, fx - 1.f, fy)));
            J1 = -1.0 * (0.5 * (double)(sample_flat(img1, lv, fx, fy + 1.f) - sample_flat(img1, lv, fx, fy - 1.f)));
        }
        // thread 0: the solver's state of this level (:44-58)
        double dx = ctl.dx, dy = ctl.dy, lastCost = 0, H00 = 0, H10 = 0, H11 = 0;
        bool succ = true;
        unsigned iters = 0;
        // thread 0, level done: the driver's hand-over to the next level (:119-123, 192-201)
        auto finish_level = [&]() {
            atomicAdd(&args.stats[kStatIters0 + level], (unsigned long long)iters);
            k2.x = kx + (float)dx;
            k2.y = ky + (float)dy;
            flag = succ && point_in_image(k2.x, k2.y, lv);
            if (level > 0) {
                k1.x = (float)((double)k1.x / 0.5);
                k1.y = (float)((double)k1.y / 0.5);
                if (flag) {
                    k2.x = (float)((double)k2.x / 0.5);
                    k2.y = (float)((double)k2.y / 0.5);
                } else {
                    k2 = k1;
                }
                ctl.kx = k1.x;
                ctl.ky = k1.y;
                ctl.dx = (double)(k2.x - k1.x);   // (levels below the top always start from the guess, :185-189)
                ctl.dy = (double)(k2.y - k1.y);
            } else if (!point_in_image(k2.x, k2.y, lv)) {
                atomicAdd(&args.stats[kStatOutOfImage], 1ull);
            }
            ctl.cmd = 0;
            ctl.level = level - 1;
        };
        if (args.max_iters <= 0) {   // no pass at all: the level keeps its initial displacement
            __syncthreads();         // (every thread has read ctl)
            if (tid == 0) finish_level();
            __syncthreads();
        } else
        for (int iter = 0;; ++iter) {
            // ---- every pixel: error and Jacobian at the current displacement
            const double ddx = ctl.dx, ddy = ctl.dy;
            double v[6] = {0, 0, 0, 0, 0, 0};
            if (active) {
                const double cx = (double)fx + ddx, cy = (double)fy + ddy;   // double adds
                const double error = (double)(i1 - sample_flat(img2, lv, (float)cx, (float)cy));
                if (!inverse) {   // :69-73
                    J0 = -1.0 * (0.5 * (double)(sample_flat(img2, lv, (float)(cx + 1), (float)cy) -
                                                sample_flat(img2, lv, (float)(cx - 1), (float)cy)));
                    J1 = -1.0 * (0.5 * (double)(sample_flat(img2, lv, (float)cx, (float)(cy + 1)) -
                                                sample_flat(img2, lv, (float)cx, (float)(cy - 1))));
                } else if (iter > 0) {   // the stale Jacobian (:57): the last pixel's of the first pass, for every pixel
                    J0 = ctl.j0;
                    J1 = ctl.j1;
                }
                v[0] = __dmul_rn(-error, J0);
                v[1] = __dmul_rn(-error, J1);
                v[2] = __dmul_rn(error, error);
                if (!inverse || iter == 0) {
                    v[3] = __dmul_rn(J0, J0);
                    v[4] = __dmul_rn(J1, J0);
                    v[5] = __dmul_rn(J1, J1);
                }
            }
            // ---- the six sums: warp trees, then the warps in order
#pragma unroll
            for (int k = 0; k < 6; ++k) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v[k] = __dadd_rn(v[k], __shfl_down_sync(0xffffffffu, v[k], o));
                if (lane == 0 && warp > 0) part[warp][k] = v[k];   // (warp 0's sums stay in thread 0's registers)
            }
            __syncthreads();
            // ---- thread 0: solve, decide (:92-117)
            if (tid == 0) {
                double s[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) s[k] = v[k];
                if (n_warps >= 2) {   // (the common case, 7x7 and 8x8: two warps -- six independent loads)
#pragma unroll
                    for (int k = 0; k < 6; ++k) s[k] = __dadd_rn(s[k], part[1][k]);
                }
                for (int w = 2; w < n_warps; ++w)
#pragma unroll
                    for (int k = 0; k < 6; ++k) s[k] = __dadd_rn(s[k], part[w][k]);
                ++iters;
                if (!inverse || iter == 0) {
                    H00 = s[3];
                    H10 = s[4];
                    H11 = s[5];
                }
                double u0, u1;
                ldlt2_solve(H00, H10, H11, s[0], s[1], u0, u1);
                bool go = true;
                if (not_finite(u0) || not_finite(u1)) {
                    atomicAdd(&args.stats[kStatNan], 1ull);
                    succ = false;
                    go = false;
                } else if (iter > 0 && s[2] > lastCost) {
                    go = false;
                } else {
                    dx = __dadd_rn(dx, u0);
                    dy = __dadd_rn(dy, u1);
                    lastCost = s[2];
                    succ = true;
                    if (__dadd_rn(__dmul_rn(u0, u0), __dmul_rn(u1, u1)) < args.eps_sq) go = false;   // norm < eps (:113)
                    if (iter + 1 >= args.max_iters) go = false;
                }
                if (go) {
                    ctl.dx = dx;
                    ctl.dy = dy;
                    ctl.cmd = 1;
                } else {
                    finish_level();
                }
            }
            if (inverse && iter == 0 && last_pixel) {   // (read by the others after the barrier below)
                ctl.j0 = J0;
                ctl.j1 = J1;
            }
            __syncthreads();
            if (ctl.cmd == 0) break;
        }
        level = ctl.level;
        // (ctl is rewritten by thread 0 only after the NEXT barrier: every thread has read kx, ky, dx, dy by then)
    }
    if (tid == 0) {
        args.kp2_out[i] = k2;
        args.success[i] = flag ? 1 : 0;
        if (flag) atomicAdd(&args.stats[kStatSuccess], 1ull);
    }
}

}  // namespace

bool patch_kernel_supports(const SolverArgs &args) {
    const int P = args.patch_hi - args.patch_lo + 1;
    return P >= 1 && P * P <= kPatchMaxThreads;
}

cudaError_t launch_klt_patch(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream) {
    if (args.n_total <= 0) return cudaSuccess;
    if (!patch_kernel_supports(args)) return cudaErrorInvalidValue;
    const int P = args.patch_hi - args.patch_lo + 1;
    const int threads = ((P * P + 31) / 32) * 32;
    klt_patch_kernel<<<args.n_total, threads, 0, stream>>>(pyr, args);
    note_launch();
    return cudaGetLastError();
}

}  // namespace legoklt
