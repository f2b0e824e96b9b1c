// klt_solver_warp.cu -- LEGO_KLT_KERNEL_WARP: one warp per feature, all pyramid levels fused (sm_100a).
//
// Replaces LKOpticalFlow1Layer + LKOpticalFlowTracker::calcLKOpticalFlow (src/algorithm.cpp:11-125)
// and the coarse-to-fine loop of LKOpticalFlow4Layer (:158-205).
//
// Mapping (north star (2)):
//   * one warp owns one feature for all levels; dx,dy,cost,lastCost and the decisions live in
//     registers, replicated on every lane (the butterfly reduction leaves identical bits everywhere,
//     so every lane takes the same branch without a broadcast);
//   * the img2 neighbourhood of the current estimate is staged in shared memory as a 48x24-byte tile
//     by TMA (cp.async.bulk.tensor.3d, one elected lane, mbarrier completion) and re-staged only when
//     the sample footprint drifts out of it;
//   * per Gauss-Newton pass the warp first evaluates the (P+2)^2 grid of bilinear samples ONCE
//     (the reference evaluates 5 samples per patch pixel = 245 for 7x7; the grid has 81), then each
//     lane forms error / gradient for its patch pixels from the grid;
//   * the six normal-equation sums are fp64 per-lane partials + __shfl_xor butterfly.
//
// Bit-fidelity contract: every fp32 value that enters the sums (error, gradient differences) is
// bit-identical to the reference's.  Sharing a sample between "centre of pixel x+1" and "+1 tap of
// pixel x" is only done after checking that the reference's three ways of forming that coordinate
// (float(double(float(kx+c))+dx), ...+1.0, ...-1.0) agree bitwise for every grid column and row; if
// they do not, or if any tap would touch the image border (where the reference clamps coordinates
// and addresses the buffer flat, algorithm.h:42-48), the pass runs on the exact per-pixel path.
// Only the ORDER of the fp64 additions differs from the reference (x-outer/y-inner sequential there,
// lane partials + butterfly here); all products are exact in fp64 (24-bit x 24-bit significands).
#include <cstdio>

#include "klt_kernels.h"

namespace legoklt {

struct alignas(64) WarpKernelMaps {
    CUtensorMap img2[kMaxLevels];
};

namespace {

#ifndef WARP_WPC
#define WARP_WPC 2   // measured: 2 warps per CTA balance small calls over the SMs and run the 11x11 instance 48 % faster than 8
#endif
constexpr int kWarpsPerCta = WARP_WPC;
constexpr int kGridStride = 16;   // row stride (floats) of the sample grid in shared memory
constexpr int kMaxGrid = kMaxPatch + 2;

struct alignas(128) WarpScratch {
    uint8_t win[kWinH * kWinW];            // TMA destination (128-byte aligned)
    float samp[kMaxGrid * kGridStride];    // bilinear sample grid
    int col_ix[16];                        // per grid column: int(X) - window origin
    float col_xx[16], col_omx[16];         //                  frac, 1 - frac
    int row_iy[16];
    float row_yy[16], row_omy[16];
    unsigned long long bar;                // mbarrier for the TMA tile
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Bounded wait: a TMA that never completes (bad descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, uint32_t parity, bool trap_on_timeout) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) {
            if (trap_on_timeout) __trap();
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ void tma_load_tile(void *dst, const CUtensorMap *map, int x, int y, int z,
                                              unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

// Out-of-line GetPixelValue: the generic sampler inlined at ~15 call sites made the kernel 7.7k SASS
// instructions (123 KB, far beyond the instruction cache); the fast path does not use it.
__device__ __noinline__ float sample_flat_ool(const uint8_t *img, const LevelView &lv, float x, float y) {
    return sample_flat(img, lv, x, y);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Coordinate of grid column/row g (offset c = lo-1+g from the keypoint) as the reference forms it,
// checking that every way the reference reaches this coordinate gives the same float.
//   centre of pixel c        : float(double(float(k + c)) + d)
//   +1 tap of pixel c-1      : float(double(float(k + (c-1))) + d + 1.0)
//   -1 tap of pixel c+1      : float(double(float(k + (c+1))) + d - 1.0)
__device__ __forceinline__ float grid_coord(float k, double d, int c, int lo, int hi, bool &consistent) {
    float v = 0.f;
    bool have = false;
    consistent = true;
    if (c >= lo && c <= hi) {
        v = (float)((double)(k + (float)c) + d);
        have = true;
    }
    if (c - 1 >= lo && c - 1 <= hi) {
        float t = (float)(((double)(k + (float)(c - 1)) + d) + 1.0);
        if (have && __float_as_uint(t) != __float_as_uint(v)) consistent = false;
        v = t;
        have = true;
    }
    if (c + 1 >= lo && c + 1 <= hi) {
        float t = (float)(((double)(k + (float)(c + 1)) + d) - 1.0);
        if (have && __float_as_uint(t) != __float_as_uint(v)) consistent = false;
        v = t;
    }
    return v;
}

// PT = compile-time patch width (0 = run-time): the generic version spent 15 % of its instructions on
// integer divisions by the run-time patch / grid width (profiles/r01_ncu_warp_kernel.csv).
template <int NR, int PT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
klt_warp_kernel(const __grid_constant__ PyramidView pyr, const __grid_constant__ WarpKernelMaps maps,
                const __grid_constant__ SolverArgs args) {
    __shared__ WarpScratch scratch[kWarpsPerCta];
    __shared__ unsigned long long s_stats[kStatCount];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch &ws = scratch[warp];
    if (threadIdx.x < kStatCount) s_stats[threadIdx.x] = 0ull;
    if (lane == 0) mbar_init(&ws.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int L = pyr.levels;
    const int lo = args.patch_lo, hi = args.patch_hi;
    const int P = PT ? PT : hi - lo + 1, G = P + 2, PP = P * P, GG = G * G;
    const bool inverse = args.inverse != 0;
    const double scale_top = 1.0 / (double)(1 << (L - 1));
    uint32_t tma_phase = 0;

    const int warps_total = gridDim.x * kWarpsPerCta;
    const int n_work = args.list ? *args.list_count : args.n_total;
    for (int wi = blockIdx.x * kWarpsPerCta + warp; wi < n_work; wi += warps_total) {
        const int f = args.list ? args.list[wi] : args.f0 + wi;
        const int img = f / args.n_per_pair;
        if (slot_unused(args, f, img)) {  // (warp-uniform)
            if (lane == 0) write_unused_slot(args, f);
            continue;
        }
        float2 k1 = args.kp1[f], k2 = args.kp2_init[f];
        k1.x = (float)(k1.x * scale_top);  // src/algorithm.cpp:160-169
        k1.y = (float)(k1.y * scale_top);
        k2.x = (float)(k2.x * scale_top);
        k2.y = (float)(k2.y * scale_top);
        bool flag = true;

        for (int level = L - 1; level >= 0; --level) {
            const LevelView &lv = pyr.lv[level];
            const uint8_t *img1 = lv.base[0] + (size_t)img * lv.slot;
            const uint8_t *img2 = lv.base[1] + (size_t)img * lv.slot;
            const bool has_initial = (level == L - 1) ? (args.has_initial != 0) : true;
            const float kx = k1.x, ky = k1.y;
            double dx = 0, dy = 0;
            if (has_initial) {
                dx = (double)(k2.x - kx);
                dy = (double)(k2.y - ky);
            }

            // Template patch I1 (constant over the iterations of this level), one value per owned
            // patch pixel; p = x_index * P + y_index  (x outer like the reference's loops).
            float I1[NR], gx1[NR], gy1[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int p = lane + 32 * r;
                I1[r] = 0.f;
                gx1[r] = gy1[r] = 0.f;
                if (p < PP) {
                    const float fx = kx + (float)(lo + p / P), fy = ky + (float)(lo + p % P);
                    I1[r] = sample_flat_ool(img1, lv, fx, fy);
                    if (inverse) {  // :74-80, evaluated on the first pass
                        gx1[r] = sample_flat_ool(img1, lv, fx + 1.f, fy) - sample_flat_ool(img1, lv, fx - 1.f, fy);
                        gy1[r] = sample_flat_ool(img1, lv, fx, fy + 1.f) - sample_flat_ool(img1, lv, fx, fy - 1.f);
                    }
                }
            }
            // inverse mode: J of the LAST patch pixel survives into later passes
            const int p_last = PP - 1;
            // (p_last lives in round p_last/32; pick the right register without dynamic indexing)
            float gxs = 0.f, gys = 0.f;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                if (r == p_last / 32) {
                    gxs = __shfl_sync(0xffffffffu, gx1[r], p_last & 31);
                    gys = __shfl_sync(0xffffffffu, gy1[r], p_last & 31);
                }
            }

            double cost = 0, lastCost = 0;
            bool succ = true;
            double H00 = 0, H01 = 0, H11 = 0;
            unsigned iters = 0, slow = 0;
            bool win_valid = false;
            int wx0 = 0, wy0 = 0;

            for (int iter = 0; iter < args.max_iters; ++iter) {
                ++iters;
                double sb0 = 0, sb1 = 0, sc = 0, s00 = 0, s01 = 0, s11 = 0;
                bool fast = false;
                if (!inverse) {
                    // ---- grid coordinates: lanes 0..G-1 columns, lanes 16..16+G-1 rows ----
                    const bool is_col = lane < G, is_row = lane >= 16 && lane < 16 + G;
                    const int g = is_col ? lane : lane - 16;
                    bool consistent = true;
                    float coord = 0.f;
                    int ic = 0;
                    bool interior = true;
                    if (is_col || is_row) {
                        coord = grid_coord(is_col ? kx : ky, is_col ? dx : dy, lo - 1 + g, lo, hi, consistent);
                        const int limit = is_col ? lv.cols : lv.rows;
                        // no clamp may trigger and both taps must be inside the image
                        interior = (coord >= 0.f) && (coord < (float)(limit - 1));
                        ic = (int)coord;
                    }
                    const bool usable = __all_sync(0xffffffffu, consistent && interior) && !(args.debug_flags & 1);
                    if (usable) {
                        const int ix_min = __shfl_sync(0xffffffffu, ic, 0);
                        const int iy_min = __shfl_sync(0xffffffffu, ic, 16);
                        bool covered = true;
                        if (is_col) covered = win_valid && ic >= wx0 && ic + 1 <= wx0 + kWinW - 1;
                        if (is_row) covered = win_valid && ic >= wy0 && ic + 1 <= wy0 + kWinH - 1;
                        if (!__all_sync(0xffffffffu, covered)) {
                            // (re)stage the tile centred on the current footprint
                            // TMA tile mode needs a 16-byte aligned innermost start coordinate (measured on
                            // B200: an unaligned uint8 x faults with 'illegal instruction'); floor to 16.
                            wx0 = (ix_min - (kWinW - 15 - (G + 1)) / 2) & ~15;
                            wy0 = iy_min - (kWinH - (G + 1)) / 2;
                            __syncwarp();
                            if (lane == 0) {
                                mbar_expect_tx(&ws.bar, kWinW * kWinH);
                                tma_load_tile(ws.win, &maps.img2[level], wx0, wy0, img, &ws.bar);
                            }
                            const bool arrived = mbar_wait(&ws.bar, tma_phase, !(args.debug_flags & 2));
                            if (!arrived && lane == 0) atomicAdd(&s_stats[kStatTmaTimeout], 1ull);
                            tma_phase ^= 1u;
                            win_valid = arrived;
                            covered = arrived;
                            if (is_col) covered = ic >= wx0 && ic + 1 <= wx0 + kWinW - 1;
                            if (is_row) covered = ic >= wy0 && ic + 1 <= wy0 + kWinH - 1;
                        }
                        fast = __all_sync(0xffffffffu, covered);
                    }
                    if (fast) {
                        const float fr = coord - floorf(coord);
                        if (is_col) {
                            ws.col_ix[g] = ic - wx0;
                            ws.col_xx[g] = fr;
                            ws.col_omx[g] = 1.f - fr;
                        }
                        if (is_row) {
                            ws.row_iy[g] = ic - wy0;
                            ws.row_yy[g] = fr;
                            ws.row_omy[g] = 1.f - fr;
                        }
                        __syncwarp();
                        // ---- sample grid ----
                        for (int s = lane; s < GG; s += 32) {
                            const int gr = s / G, gc = s - gr * G;
                            const uint8_t *p = ws.win + ws.row_iy[gr] * kWinW + ws.col_ix[gc];
                            const float xx = ws.col_xx[gc], omx = ws.col_omx[gc];
                            const float yy = ws.row_yy[gr], omy = ws.row_omy[gr];
                            float v = __fmul_rn(__fmul_rn(omx, omy), (float)p[0]);
                            v = __fadd_rn(v, __fmul_rn(__fmul_rn(xx, omy), (float)p[1]));
                            v = __fadd_rn(v, __fmul_rn(__fmul_rn(omx, yy), (float)p[kWinW]));
                            v = __fadd_rn(v, __fmul_rn(__fmul_rn(xx, yy), (float)p[kWinW + 1]));
                            ws.samp[gr * kGridStride + gc] = v;
                        }
                        __syncwarp();
                        // ---- patch pixels ----
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            const int p = lane + 32 * r;
                            if (p < PP) {
                                const int gc = p / P + 1, gr = p % P + 1;
                                const float *c = ws.samp + gr * kGridStride + gc;
                                const double e = (double)(I1[r] - c[0]);
                                const double gx = (double)(c[1] - c[-1]);
                                const double gy = (double)(c[kGridStride] - c[-kGridStride]);
                                sb0 = fma(e, gx, sb0);
                                sb1 = fma(e, gy, sb1);
                                sc = fma(e, e, sc);
                                s00 = fma(gx, gx, s00);
                                s01 = fma(gx, gy, s01);
                                s11 = fma(gy, gy, s11);
                            }
                        }
                        __syncwarp();
                    } else {
                        // ---- exact per-pixel path (borders, inconsistent coordinates) ----
                        ++slow;
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            const int p = lane + 32 * r;
                            if (p < PP) {
                                const float fx = kx + (float)(lo + p / P), fy = ky + (float)(lo + p % P);
                                const double cx = (double)fx + dx, cy = (double)fy + dy;
                                const double e = (double)(I1[r] - sample_flat_ool(img2, lv, (float)cx, (float)cy));
                                const double gx = (double)(sample_flat_ool(img2, lv, (float)(cx + 1), (float)cy) -
                                                           sample_flat_ool(img2, lv, (float)(cx - 1), (float)cy));
                                const double gy = (double)(sample_flat_ool(img2, lv, (float)cx, (float)(cy + 1)) -
                                                           sample_flat_ool(img2, lv, (float)cx, (float)(cy - 1)));
                                sb0 = fma(e, gx, sb0);
                                sb1 = fma(e, gy, sb1);
                                sc = fma(e, e, sc);
                                s00 = fma(gx, gx, s00);
                                s01 = fma(gx, gy, s01);
                                s11 = fma(gy, gy, s11);
                            }
                        }
                    }
                } else {
                    // ---- the reference's inverse mode: J from img1 on pass 0, stale afterwards ----
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        const int p = lane + 32 * r;
                        if (p < PP) {
                            const float fx = kx + (float)(lo + p / P), fy = ky + (float)(lo + p % P);
                            const double cx = (double)fx + dx, cy = (double)fy + dy;
                            const double e = (double)(I1[r] - sample_flat_ool(img2, lv, (float)cx, (float)cy));
                            const double gx = (double)(iter == 0 ? gx1[r] : gxs);
                            const double gy = (double)(iter == 0 ? gy1[r] : gys);
                            sb0 = fma(e, gx, sb0);
                            sb1 = fma(e, gy, sb1);
                            sc = fma(e, e, sc);
                            if (iter == 0) {
                                s00 = fma(gx, gx, s00);
                                s01 = fma(gx, gy, s01);
                                s11 = fma(gy, gy, s11);
                            }
                        }
                    }
                }

                // ---- reduce, rescale (J = -0.5*g: exact powers of two), solve, decide ----
                const double b0 = 0.5 * warp_sum(sb0);
                const double b1 = 0.5 * warp_sum(sb1);
                cost = warp_sum(sc);
                if (!inverse || iter == 0) {
                    H00 = 0.25 * warp_sum(s00);
                    H01 = 0.25 * warp_sum(s01);
                    H11 = 0.25 * warp_sum(s11);
                }
                double u0, u1;
                ldlt2_solve(H00, H01, H11, b0, b1, u0, u1);
                if (not_finite(u0) || not_finite(u1)) {  // :94-100
                    if (lane == 0) atomicAdd(&s_stats[kStatNan], 1ull);
                    succ = false;
                    break;
                }
                if (iter > 0 && cost > lastCost) break;  // :102-104
                dx = __dadd_rn(dx, u0);                  // :107-110
                dy = __dadd_rn(dy, u1);
                lastCost = cost;
                succ = true;
                if (__dadd_rn(__dmul_rn(u0, u0), __dmul_rn(u1, u1)) < args.eps_sq) break;  // :113
            }

            if (lane == 0) {
                atomicAdd(&s_stats[kStatIters0 + level], (unsigned long long)iters);
                if (slow) atomicAdd(&s_stats[kStatSlowPath], (unsigned long long)slow);
            }
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
            } else if (lane == 0 && !point_in_image(k2.x, k2.y, lv)) {
                atomicAdd(&s_stats[kStatOutOfImage], 1ull);
            }
        }
        if (lane == 0) {
            args.kp2_out[f] = k2;
            args.success[f] = flag ? 1 : 0;
            if (flag) atomicAdd(&s_stats[kStatSuccess], 1ull);
        }
    }
    __syncthreads();
    if (threadIdx.x < kStatCount && s_stats[threadIdx.x]) atomicAdd(&args.stats[threadIdx.x], s_stats[threadIdx.x]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

cudaError_t warp_maps_create(const PyramidView &pyr, WarpKernelMaps **out) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return cudaErrorNotSupported;
    WarpKernelMaps *m = new WarpKernelMaps();
    memset(m, 0, sizeof(*m));
    for (int l = 0; l < pyr.levels; ++l) {
        const LevelView &lv = pyr.lv[l];
        cuuint64_t dims[3] = {(cuuint64_t)lv.cols, (cuuint64_t)lv.rows, (cuuint64_t)pyr.n_images};
        cuuint64_t strides[2] = {(cuuint64_t)lv.pitch, (cuuint64_t)lv.slot};
        cuuint32_t box[3] = {(cuuint32_t)kWinW, (cuuint32_t)kWinH, 1u};
        cuuint32_t estr[3] = {1u, 1u, 1u};
        CUresult rc = enc(&m->img2[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, lv.base[1], dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) {
            delete m;
            return cudaErrorInvalidValue;
        }
    }
    *out = m;
    return cudaSuccess;
}

void warp_maps_destroy(WarpKernelMaps *maps) { delete maps; }

cudaError_t launch_klt_warp(const PyramidView &pyr, const WarpKernelMaps *maps, const SolverArgs &args,
                            int sm_count, cudaStream_t stream) {
    if (args.n_total <= 0) return cudaSuccess;
    const int P = args.patch_hi - args.patch_lo + 1;
    const int PP = P * P;
    const int block = kWarpsPerCta * 32;
    int ctas_needed = (args.n_total + kWarpsPerCta - 1) / kWarpsPerCta;
    int grid = sm_count * (64 / kWarpsPerCta);  // persistent warps (64 per SM launched), grid-stride over features
    if (grid > ctas_needed) grid = ctas_needed;
    if (P == 7)
        klt_warp_kernel<2, 7><<<grid, block, 0, stream>>>(pyr, *maps, args);
    else if (P == 8)
        klt_warp_kernel<2, 8><<<grid, block, 0, stream>>>(pyr, *maps, args);
    else if (P == 11)
        klt_warp_kernel<4, 11><<<grid, block, 0, stream>>>(pyr, *maps, args);
    else if (PP <= 64)
        klt_warp_kernel<2, 0><<<grid, block, 0, stream>>>(pyr, *maps, args);
    else if (PP <= 128)
        klt_warp_kernel<4, 0><<<grid, block, 0, stream>>>(pyr, *maps, args);
    else
        klt_warp_kernel<6, 0><<<grid, block, 0, stream>>>(pyr, *maps, args);
    note_launch();
    return cudaGetLastError();
}

}  // namespace legoklt
