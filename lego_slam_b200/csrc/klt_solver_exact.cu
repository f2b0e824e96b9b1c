// klt_solver_exact.cu -- LEGO_KLT_KERNEL_EXACT: one thread per feature, reference operation order.
//
// Restates src/algorithm.cpp:37-125 (solver) and :158-205 (coarse-to-fine driver) with every level
// fused into one kernel: features are independent, so no grid-wide barrier is needed between levels
// (SURVEY.md 3.2).  Accumulation order (x outer, y inner, sequential fp64) and every rounding step
// are the reference's, so results are bit-identical to the CPU oracle.  It is the on-GPU checker for
// the fast kernels and the path for patch shapes the fast kernels do not cover; it is NOT fast.
#include "klt_common.cuh"
#include "klt_kernels.h"

namespace legoklt {

__global__ void __launch_bounds__(128)
klt_exact_kernel(const __grid_constant__ PyramidView pyr, const __grid_constant__ SolverArgs args) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= args.n_total) return;
    const int i = args.f0 + t;
    const int img = i / args.n_per_pair;
    if (slot_unused(args, i, img)) {
        write_unused_slot(args, i);
        return;
    }
    const int L = pyr.levels;
    const bool inverse = args.inverse != 0;

    // :160-169  scale to the top level: Point2f *= double
    const double scale_top = 1.0 / (double)(1 << (L - 1));
    float2 k1 = args.kp1[i], k2 = args.kp2_init[i];
    k1.x = (float)(k1.x * scale_top);
    k1.y = (float)(k1.y * scale_top);
    k2.x = (float)(k2.x * scale_top);
    k2.y = (float)(k2.y * scale_top);
    bool flag = true;

    for (int level = L - 1; level >= 0; --level) {
        const LevelView &lv = pyr.lv[level];
        const uint8_t *img1 = lv.base[0] + (size_t)img * lv.slot;
        const uint8_t *img2 = lv.base[1] + (size_t)img * lv.slot;
        const bool has_initial = (level == L - 1) ? (args.has_initial != 0) : true;  // :185-189
        const float kx = k1.x, ky = k1.y;
        double dx = 0, dy = 0;
        if (has_initial) {  // :47-50
            dx = (double)(k2.x - kx);
            dy = (double)(k2.y - ky);
        }
        double cost = 0, lastCost = 0;
        bool succ = true;
        double H00 = 0, H01 = 0, H10 = 0, H11 = 0, b0 = 0, b1 = 0, J0 = 0, J1 = 0;
        unsigned iters = 0;
        for (int iter = 0; iter < args.max_iters; ++iter) {
            if (!inverse) { H00 = H01 = H10 = H11 = 0; }
            b0 = b1 = 0;
            cost = 0;
            ++iters;
            for (int x = args.patch_lo; x <= args.patch_hi; ++x) {
                for (int y = args.patch_lo; y <= args.patch_hi; ++y) {
                    const float fx = kx + (float)x, fy = ky + (float)y;      // float adds
                    const double cx = (double)fx + dx, cy = (double)fy + dy;  // double adds
                    double error = (double)(sample_flat(img1, lv, fx, fy) -
                                            sample_flat(img2, lv, (float)cx, (float)cy));
                    if (!inverse) {
                        J0 = -1.0 * (0.5 * (double)(sample_flat(img2, lv, (float)(cx + 1), (float)cy) -
                                                    sample_flat(img2, lv, (float)(cx - 1), (float)cy)));
                        J1 = -1.0 * (0.5 * (double)(sample_flat(img2, lv, (float)cx, (float)(cy + 1)) -
                                                    sample_flat(img2, lv, (float)cx, (float)(cy - 1))));
                    } else if (iter == 0) {
                        J0 = -1.0 * (0.5 * (double)(sample_flat(img1, lv, fx + 1.f, fy) -
                                                    sample_flat(img1, lv, fx - 1.f, fy)));
                        J1 = -1.0 * (0.5 * (double)(sample_flat(img1, lv, fx, fy + 1.f) -
                                                    sample_flat(img1, lv, fx, fy - 1.f)));
                    }
                    b0 = __dadd_rn(b0, __dmul_rn(-error, J0));
                    b1 = __dadd_rn(b1, __dmul_rn(-error, J1));
                    cost = __dadd_rn(cost, __dmul_rn(error, error));
                    if (!inverse || iter == 0) {
                        H00 = __dadd_rn(H00, __dmul_rn(J0, J0));
                        H01 = __dadd_rn(H01, __dmul_rn(J0, J1));
                        H10 = __dadd_rn(H10, __dmul_rn(J1, J0));
                        H11 = __dadd_rn(H11, __dmul_rn(J1, J1));
                    }
                }
            }
            double u0, u1;
            ldlt2_solve(H00, H10, H11, b0, b1, u0, u1);
            if (not_finite(u0) || not_finite(u1)) {
                atomicAdd(&args.stats[kStatNan], 1ull);
                succ = false;
                break;
            }
            if (iter > 0 && cost > lastCost) break;
            dx = __dadd_rn(dx, u0);
            dy = __dadd_rn(dy, u1);
            lastCost = cost;
            succ = true;
            if (sqrt(__dadd_rn(__dmul_rn(u0, u0), __dmul_rn(u1, u1))) < args.eps) break;
        }
        atomicAdd(&args.stats[kStatIters0 + level], (unsigned long long)iters);
        k2.x = kx + (float)dx;  // :121
        k2.y = ky + (float)dy;
        flag = succ && point_in_image(k2.x, k2.y, lv);  // :119,123
        if (level > 0) {  // :192-201
            k1.x = (float)((double)k1.x / 0.5);
            k1.y = (float)((double)k1.y / 0.5);
            if (flag) {
                k2.x = (float)((double)k2.x / 0.5);
                k2.y = (float)((double)k2.y / 0.5);
            } else {
                k2 = k1;
            }
        } else if (!point_in_image(k2.x, k2.y, lv)) {
            atomicAdd(&args.stats[kStatOutOfImage], 1ull);
        }
    }
    args.kp2_out[i] = k2;
    args.success[i] = flag ? 1 : 0;
    if (flag) atomicAdd(&args.stats[kStatSuccess], 1ull);
}

cudaError_t launch_klt_exact(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream) {
    if (args.n_total <= 0) return cudaSuccess;
    int block = 128;
    int grid = (args.n_total + block - 1) / block;
    klt_exact_kernel<<<grid, block, 0, stream>>>(pyr, args);
    note_launch();
    return cudaGetLastError();
}

}  // namespace legoklt
