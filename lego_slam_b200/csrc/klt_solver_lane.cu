// klt_solver_lane.cu -- LEGO_KLT_KERNEL_LANE: one THREAD per feature, persistent per-thread state
// machine, all pyramid levels fused (sm_100a).  7x7 patch, forward mode (the reference's call-site
// configuration: src/frontend_g2o.cpp:473,515 with src/algorithm.cpp:40 half_patch_size = 3).
//
// Replaces LKOpticalFlow1Layer + LKOpticalFlowTracker::calcLKOpticalFlow (src/algorithm.cpp:11-125)
// and the coarse-to-fine loop of LKOpticalFlow4Layer (:158-205).
//
// Why not warp-per-feature (measured, profiles/r01_warp_kernel.md): that mapping spends ~850 warp
// instructions per Gauss-Newton pass -- shuffles for six fp64 reductions, a serial 2x2 solve on every
// lane, per-sample index arithmetic -- for 49 pixels of useful work.  Here every lane runs a whole
// pass for its own feature (no shuffles, no idle lanes in the solve), ~70 warp instructions per
// feature-pass.  Divergence in iteration counts (1..10 per level) is removed by a state machine: each
// trip of the main loop is exactly one pass for every live thread; a thread that converges moves to
// its next level / next feature while its neighbours keep iterating.
//
// Data movement: global memory is never read divergently per thread.  When a thread starts a level
// (or its footprint drifts out of its window) the WARP stages for it, cooperatively and coalesced:
//   * the 7x7 template patch I1 -- sampled with the reference's exact border semantics from a staged
//     img1 tile, stored as 49 floats in the thread's shared-memory column;
//   * a 32x14-byte img2 window (16-byte aligned origin, replicate-clamped at the image border) in a
//     word-interleaved layout  win2[word][T+1]  -> bank = (word + thread) % 32: conflict-free for the
//     cooperative writer (fixed thread, consecutive words) and for the pass (fixed word, all threads).
//
// Bit-fidelity contract (same as the warp kernel): every fp32 value entering the sums is bit-identical
// to the reference's; only the ORDER of the fp64 additions differs (row-major here, x-outer there; all
// products are exact in fp64).  The (P+2)^2 sample grid is shared between "centre of pixel x+1" and
// "+1 tap of pixel x" only when that is provably what the reference computes:
//   (a) kx+c, ky+c are exact in fp32 for c in [-3,3] (checked per level);
//   (b) the double coordinate is not within 16 ulp64 of an fp32 rounding midpoint (checked per pass,
//       per grid column/row), so the <=2 ulp64 differences between the reference's three ways of
//       forming a coordinate cannot change the rounded float -- or it is EXACTLY on a midpoint and all
//       double sums are provably exact (TwoSum), which is the common first-pass case (kx + float dx);
// The border semantics of algorithm.h:42-55 ARE part of the fast path: clamped coordinates get the
// factors (1,0) over replicated border pixels; a sample in the last-pixel sliver x in (cols-1, cols)
// reads the reference's flat-address neighbour (first pixel of the next row), which the stager puts in
// window column `cols`; y in (rows-1, rows) reads zeros below the image (the oracle's definition of the
// reference's out-of-buffer read).  A feature that violates (a) or (b) is appended to a deferred list
// and finished by the exact warp kernel; its partial work is dropped.
#include "klt_kernels.h"

namespace legoklt {

namespace {

constexpr int LO = -3, HI = 3, P = 7, G = 9;
constexpr int kWin2Rows = 14, kWin2Words = 8;                 // 14 rows x 32 bytes per thread
constexpr int kWin2Total = kWin2Rows * kWin2Words;            // 112 words
constexpr int kI1Count = P * P;                               // 49 floats
constexpr int kScratchRows = 8, kScratchW = 32;               // img1 staging tile per warp

enum : int { ST_FETCH = 0, ST_LEVEL = 1, ST_RUN = 2, ST_DONE = 3 };

template <int T>
struct LaneSmem {
    uint32_t win2[kWin2Total][T + 1];
    float i1[kI1Count][T + 1];
    alignas(16) uint8_t scratch[T / 32][kScratchRows * kScratchW];
    unsigned stats[kStatCount];
};

__device__ __forceinline__ float byte_to_float(uint32_t packed, int k) {
    // place byte k in the low mantissa of 2^23 and subtract 2^23: exact, ALU + FADD, no XU convert
    return __fadd_rn(__uint_as_float(__byte_perm(packed, 0x4B000000u, 0x7440u + k)), -8388608.0f);
}

// Reference bilinear formula (algorithm.h:51-56), per-sample weights, individually rounded ops.
__device__ __forceinline__ float bilerp(float omx, float xx, float omy, float yy, float p0, float p1, float p2,
                                        float p3) {
    float r = __fmul_rn(__fmul_rn(omx, omy), p0);
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xx, omy), p1));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(omx, yy), p2));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(xx, yy), p3));
    return r;
}

// One grid axis (columns or rows) of the shared sample grid.  S = (double)k + d; the reference forms
// float(double(float(k+c)) + d [+-1]); under condition (a) these are S + g - 4 up to 2 ulp64.
// Per grid index g it yields the two bilinear factors the reference would use:
//     om[g] = 1 - frac   (weight of the tap at the integer coordinate)
//     fr[g] = frac       (weight of the tap at integer coordinate + 1)
// with the nominal integer coordinate of index g being origin + g (consecutive), and the border
// semantics of algorithm.h:42-55 folded into the factors:
//   * X <  0      : clamped to 0        -> (1, 0); the window holds replicated border pixels
//   * X >= limit  : clamped to limit-1  -> (1, 0), or (0, 1) when origin+g == limit exactly, because
//                   window column `limit` holds the flat-address "wrap" pixel (see the stager) and the
//                   replicated border pixel sits one further
//   * sliver X in (limit-1, limit), not clamped: the +1 tap is data[.. + 1] on the next row (columns:
//     the wrap pixel, true factors) or past the buffer = 0 (rows: fr := 0, om = 1 - frac).
// Returns 0, or the reason (kStatDefer*) the pass cannot be proven bit-identical on this grid.
template <bool IS_ROW>
__device__ __forceinline__ int grid_axis(float k, double d, int limit, int &origin, float (&fr)[G], float (&om)[G]) {
    const double kd = (double)k;
    const double S = kd + d;
    if (!(fabs(S) < 1.0e6)) return kStatDeferRange;
    // An EXACT tie (D on an fp32 rounding midpoint) is safe when every double sum involved is exact:
    // then all of the reference's ways of forming the coordinate give the same double, and round-half-
    // even gives the same float.  TwoSum error of S, and 4 spare low bits so that S + c stays exact.
    const double bb = S - kd;
    const double err = (kd - (S - bb)) + (d - bb);
    const bool tie_ok = (err == 0.0) && ((__double2loint(S) & 0xF) == 0) && (fabs(S) >= 1.0);
    const double D0 = S + (double)(LO - 1);
    origin = __double2int_rd(D0);  // nominal integer coordinate of grid index 0
    int why = 0;  // 0 = regular, else the kStatDefer* reason
    const float flimit = (float)limit, flast = (float)(limit - 1);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const double D = S + (double)(LO - 1 + g);
        const float X = (float)D;
        // (b) distance of D's discarded mantissa bits from the fp32 rounding midpoint
        const int low = __double2loint(D) & 0x1FFFFFFF;
        const int dist = low - 0x10000000;
        if (dist == 0 ? !tie_ok : (abs(dist) <= 16)) why = kStatDeferMargin;
        const bool clamp_lo = X < 0.f, clamp_hi = X >= flimit;
        float f = __fadd_rn(X, -(float)(origin + g));
        float o = __fadd_rn(1.f, -f);
        if (clamp_lo) {
            f = 0.f;
            o = 1.f;
        } else if (clamp_hi) {
            const bool at_wrap = !IS_ROW && (origin + g == limit);
            f = at_wrap ? 1.f : 0.f;
            o = at_wrap ? 0.f : 1.f;
        } else {
            if (!((f >= 0.f) && (f <= 1.f))) why = kStatDeferNominal;
            if (IS_ROW && X > flast) f = 0.f;  // taps below the last row read zeros
        }
        fr[g] = f;
        om[g] = o;
    }
    return why;
}

// GetPixelValue on img1 for the cooperative template stage: taps from the warp's staged tile when all
// four are inside it and inside the image, else the reference's flat addressing from global memory.
__device__ __forceinline__ float sample_staged(const uint8_t *tile, int sx0, int sy0, const uint8_t *img,
                                               const LevelView &lv, float x, float y) {
    if (x < 0.f) x = 0.f;
    if (y < 0.f) y = 0.f;
    if (x >= (float)lv.cols) x = (float)(lv.cols - 1);
    if (y >= (float)lv.rows) y = (float)(lv.rows - 1);
    const int ix = (int)x, iy = (int)y;
    const float xx = x - floorf(x), yy = y - floorf(y);
    float p0, p1, p2, p3;
    const int tx = ix - sx0, ty = iy - sy0;
    if (ix + 1 < lv.cols && iy + 1 < lv.rows && tx >= 0 && tx + 1 < kScratchW && ty >= 0 && ty + 1 < kScratchRows) {
        const uint8_t *p = tile + ty * kScratchW + tx;
        p0 = (float)p[0];
        p1 = (float)p[1];
        p2 = (float)p[kScratchW];
        p3 = (float)p[kScratchW + 1];
    } else if (ix + 1 < lv.cols && iy + 1 < lv.rows) {
        const uint8_t *p = img + (size_t)iy * lv.pitch + ix;
        p0 = (float)__ldg(p);
        p1 = (float)__ldg(p + 1);
        p2 = (float)__ldg(p + lv.pitch);
        p3 = (float)__ldg(p + lv.pitch + 1);
    } else {
        const long long f = (long long)iy * lv.step + ix;
        p0 = fetch_flat(img, lv, f);
        p1 = fetch_flat(img, lv, f + 1);
        p2 = fetch_flat(img, lv, f + lv.step);
        p3 = fetch_flat(img, lv, f + lv.step + 1);
    }
    return bilerp(1.f - xx, xx, 1.f - yy, yy, p0, p1, p2, p3);
}

template <int T, int MIN_CTAS>
__global__ void __launch_bounds__(T, MIN_CTAS)
klt_lane_kernel(const __grid_constant__ PyramidView pyr, const __grid_constant__ SolverArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LaneSmem<T> &sm = *reinterpret_cast<LaneSmem<T> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int WS = T + 1;  // word stride between consecutive window words of one thread

    if (tid < kStatCount) sm.stats[tid] = 0u;
    __syncthreads();

    const int L = pyr.levels;
    const double scale_top = 1.0 / (double)(1 << (L - 1));
    uint8_t *tile = sm.scratch[warp];

    // ---- per-thread (per-feature) state ----
    int state = ST_FETCH, feat = -1, level = 0, iter = 0;
    float2 k1 = make_float2(0.f, 0.f), k2 = make_float2(0.f, 0.f);
    float kx = 0.f, ky = 0.f;
    double dx = 0, dy = 0, lastCost = 0;
    bool succ = true, flag = true;
    unsigned iters_packed = 0, nan_count = 0;
    int img = 0;
    int wx0 = 0, wy0 = 0;
    bool win_ok = false, need_i1 = false;

    for (;;) {
        // ------------------------------------------------------------------ fetch new features
        {
            const unsigned m = __ballot_sync(FULL, state == ST_FETCH);
            if (m) {
                const int leader = __ffs(m) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(args.work_counter, __popc(m));
                base = __shfl_sync(FULL, base, leader);
                if (state == ST_FETCH) {
                    const int id = base + __popc(m & ((1u << lane) - 1u));
                    if (id < args.n_total) {
                        feat = id;
                        img = id / args.n_per_pair;
                        k1 = args.kp1[id];
                        k2 = args.kp2_init[id];
                        k1.x = (float)(k1.x * scale_top);  // src/algorithm.cpp:160-169
                        k1.y = (float)(k1.y * scale_top);
                        k2.x = (float)(k2.x * scale_top);
                        k2.y = (float)(k2.y * scale_top);
                        level = L - 1;
                        iters_packed = 0;
                        nan_count = 0;
                        flag = true;
                        state = ST_LEVEL;
                    } else {
                        state = ST_DONE;
                    }
                }
            }
            if (__all_sync(FULL, state == ST_DONE)) break;
        }

        bool defer = false;
        int defer_why = kStatDeferInexact;

        // ------------------------------------------------------------------ level setup (scalar)
        if (state == ST_LEVEL) {
            const bool has_initial = (level == L - 1) ? (args.has_initial != 0) : true;  // :185-189
            kx = k1.x;
            ky = k1.y;
            dx = dy = 0;
            if (has_initial) {  // :47-50
                dx = (double)(k2.x - kx);
                dy = (double)(k2.y - ky);
            }
            iter = 0;
            lastCost = 0;
            succ = true;
            win_ok = false;
            need_i1 = true;
            // (a) kx+c, ky+c exact for c in [LO,HI]: the end of larger magnitude decides
            const double kxd = (double)kx, kyd = (double)ky;
            const bool exact = ((double)(kx + (float)LO) == kxd + (double)LO) && ((double)(kx + (float)HI) == kxd + (double)HI) &&
                               ((double)(ky + (float)LO) == kyd + (double)LO) && ((double)(ky + (float)HI) == kyd + (double)HI);
            if (!exact) defer = true;
            state = ST_RUN;
        }

        // ------------------------------------------------------------------ grid coordinates
        const bool live = (state == ST_RUN) && !defer;
        float xx[G], omx[G], yy[G], omy[G];
        int ixn = 0, iyn = 0;
        bool need_win = false;
        if (live) {
            const LevelView &lv = pyr.lv[level];
            const int whyx = grid_axis<false>(kx, dx, lv.cols, ixn, xx, omx);
            const int whyy = grid_axis<true>(ky, dy, lv.rows, iyn, yy, omy);
            if (whyx | whyy) {
                defer = true;
                defer_why = whyx ? whyx : whyy;
            } else {
                const bool covered = win_ok && ixn >= wx0 && ixn + (G + 1) <= wx0 + kWin2Words * 4 && iyn >= wy0 &&
                                     iyn + (G + 1) <= wy0 + kWin2Rows;
                if (!covered) {
                    wx0 = (ixn - 2) & ~15;
                    wy0 = iyn - 2;
                    need_win = true;
                    win_ok = true;
                }
            }
        }

        if (defer) {
            const int di = atomicAdd(args.defer_count, 1);
            args.defer_list[di] = feat;
            atomicAdd(&sm.stats[kStatDeferred], 1u);
            atomicAdd(&sm.stats[defer_why], 1u);
            state = ST_FETCH;
            need_i1 = false;
        }
        const bool run = (state == ST_RUN);

        // ------------------------------------------------------------------ cooperative staging
        {
            unsigned pending = __ballot_sync(FULL, run && (need_i1 || need_win));
            while (pending) {
                const int j = __ffs(pending) - 1;
                pending &= pending - 1;
                const int j_level = __shfl_sync(FULL, level, j);
                const int j_img = __shfl_sync(FULL, img, j);
                const int j_flags = __shfl_sync(FULL, (need_i1 ? 1 : 0) | (need_win ? 2 : 0), j);
                const LevelView &lv = pyr.lv[j_level];
                const int jt = (tid & ~31) + j;  // the staged thread's column in the interleaved arrays
                if (j_flags & 1) {
                    const float jkx = __shfl_sync(FULL, kx, j), jky = __shfl_sync(FULL, ky, j);
                    const uint8_t *img1 = lv.base[0] + (size_t)j_img * lv.slot;
                    const int sx0 = max(0, (int)floorf(jkx + (float)LO)) & ~15;
                    const int sy0 = max(0, (int)floorf(jky + (float)LO));
                    if (lane < 2 * kScratchRows) {
                        const int row = sy0 + (lane >> 1), qx = sx0 + 16 * (lane & 1);
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (row < lv.rows && qx + 16 <= lv.pitch)
                            v = __ldg(reinterpret_cast<const uint4 *>(img1 + (size_t)row * lv.pitch + qx));
                        *reinterpret_cast<uint4 *>(tile + (lane >> 1) * kScratchW + 16 * (lane & 1)) = v;
                    }
                    __syncwarp();
                    for (int p = lane; p < kI1Count; p += 32) {
                        const float fx = jkx + (float)(LO + p / P), fy = jky + (float)(LO + p % P);  // :65
                        sm.i1[p][jt] = sample_staged(tile, sx0, sy0, img1, lv, fx, fy);
                    }
                    __syncwarp();
                }
                if (j_flags & 2) {
                    const int jwx = __shfl_sync(FULL, wx0, j), jwy = __shfl_sync(FULL, wy0, j);
                    const uint8_t *img2 = lv.base[1] + (size_t)j_img * lv.slot;
                    if (lane < 2 * kWin2Rows) {
                        const int row = lane >> 1, half = lane & 1;
                        const int ry = min(max(jwy + row, 0), lv.rows - 1);
                        const int qx = jwx + 16 * half;
                        const uint8_t *rp = img2 + (size_t)ry * lv.pitch;
                        uint4 v;
                        if (qx >= 0 && qx + 16 <= lv.cols) {
                            v = __ldg(reinterpret_cast<const uint4 *>(rp + qx));
                        } else {  // image border: replicate left/right, flat-address wrap pixel at column `cols`
                            uint32_t w[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint32_t acc = 0;
#pragma unroll
                                for (int b = 0; b < 4; ++b) {
                                    const int n = qx + 4 * q + b;
                                    uint32_t px;
                                    if (n == lv.cols)  // data[ry*step + (cols-1) + 1] of algorithm.h:48,53
                                        px = (uint32_t)fetch_flat(img2, lv, (long long)ry * lv.step + lv.cols);
                                    else
                                        px = (uint32_t)__ldg(rp + min(max(n, 0), lv.cols - 1));
                                    acc |= px << (8 * b);
                                }
                                w[q] = acc;
                            }
                            v = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                        uint32_t *dst = &sm.win2[row * kWin2Words + half * 4][jt];
                        dst[0] = v.x;
                        dst[WS] = v.y;
                        dst[2 * WS] = v.z;
                        dst[3 * WS] = v.w;
                    }
                }
                __syncwarp();
            }
            need_i1 = false;
        }

        // ------------------------------------------------------------------ one Gauss-Newton pass
        if (run) {
            const LevelView &lv = pyr.lv[level];
            const int ox = ixn - wx0;
            const int sh = (ox & 3) * 8;
            const uint32_t *wp = &sm.win2[(iyn - wy0) * kWin2Words + (ox >> 2)][tid];
            const float *i1p = &sm.i1[0][tid];

            double sb0 = 0, sb1 = 0, sc = 0, s00 = 0, s01 = 0, s11 = 0;
            float rowA[G + 1], rowB[G + 1];
            float S[3][G];

            auto load_row = [&](int i, float (&row)[G + 1]) {
                const uint32_t *p = wp + i * kWin2Words * WS;
                const uint32_t w0 = p[0], w1 = p[WS], w2 = p[2 * WS], w3 = p[3 * WS];
                const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh),
                               b2 = __funnelshift_r(w2, w3, sh);
                row[0] = byte_to_float(b0, 0);
                row[1] = byte_to_float(b0, 1);
                row[2] = byte_to_float(b0, 2);
                row[3] = byte_to_float(b0, 3);
                row[4] = byte_to_float(b1, 0);
                row[5] = byte_to_float(b1, 1);
                row[6] = byte_to_float(b1, 2);
                row[7] = byte_to_float(b1, 3);
                row[8] = byte_to_float(b2, 0);
                row[9] = byte_to_float(b2, 1);
            };

            load_row(0, rowA);
#pragma unroll
            for (int r = 0; r < G; ++r) {
                load_row(r + 1, rowB);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    // the four corner samples of the grid are never used
                    if ((r == 0 || r == G - 1) && (g == 0 || g == G - 1)) continue;
                    S[r % 3][g] = bilerp(omx[g], xx[g], omy[r], yy[r], rowA[g], rowA[g + 1], rowB[g], rowB[g + 1]);
                }
#pragma unroll
                for (int g = 0; g <= G; ++g) rowA[g] = rowB[g];
                if (r >= 2) {
                    const int y = r - 2;  // patch row (offset LO + y), centre samples live in grid row r-1
#pragma unroll
                    for (int x = 0; x < P; ++x) {
                        const int g = x + 1;
                        const float i1v = i1p[(x * P + y) * WS];
                        const double e = (double)__fadd_rn(i1v, -S[(r - 1) % 3][g]);                       // :65-66
                        const double gx = (double)__fadd_rn(S[(r - 1) % 3][g + 1], -S[(r - 1) % 3][g - 1]);  // :70-71
                        const double gy = (double)__fadd_rn(S[r % 3][g], -S[(r - 2) % 3][g]);              // :72-73
                        sb0 = fma(e, gx, sb0);
                        sb1 = fma(e, gy, sb1);
                        sc = fma(e, e, sc);
                        s00 = fma(gx, gx, s00);
                        s01 = fma(gx, gy, s01);
                        s11 = fma(gy, gy, s11);
                    }
                }
            }

            // J = -0.5 * g: rescaling the sums by exact powers of two commutes with every rounding
            const double b0 = 0.5 * sb0, b1 = 0.5 * sb1, cost = sc;
            const double H00 = 0.25 * s00, H01 = 0.25 * s01, H11 = 0.25 * s11;
            ++iter;
            bool level_done = false;
            double u0, u1;
            ldlt2_solve(H00, H01, H11, b0, b1, u0, u1);          // :92-93
            if (not_finite(u0) || not_finite(u1)) {               // :94-100
                ++nan_count;
                succ = false;
                level_done = true;
            } else if (iter > 1 && cost > lastCost) {             // :102-104 (iter here is 1-based)
                level_done = true;
            } else {
                dx = __dadd_rn(dx, u0);                           // :107-110
                dy = __dadd_rn(dy, u1);
                lastCost = cost;
                succ = true;
                if (sqrt(__dadd_rn(__dmul_rn(u0, u0), __dmul_rn(u1, u1))) < args.eps) level_done = true;  // :113
                if (iter >= args.max_iters) level_done = true;
            }

            if (level_done) {
                iters_packed |= (unsigned)iter << (4 * level);
                k2.x = kx + (float)dx;                            // :121
                k2.y = ky + (float)dy;
                flag = succ && point_in_image(k2.x, k2.y, lv);    // :119,123
                if (level > 0) {                                  // :192-201
                    k1.x = (float)((double)k1.x / 0.5);
                    k1.y = (float)((double)k1.y / 0.5);
                    if (flag) {
                        k2.x = (float)((double)k2.x / 0.5);
                        k2.y = (float)((double)k2.y / 0.5);
                    } else {
                        k2 = k1;
                    }
                    --level;
                    state = ST_LEVEL;
                } else {
                    args.kp2_out[feat] = k2;
                    args.success[feat] = flag ? 1 : 0;
                    for (int l = 0; l < L; ++l) atomicAdd(&sm.stats[kStatIters0 + l], (iters_packed >> (4 * l)) & 15u);
                    if (nan_count) atomicAdd(&sm.stats[kStatNan], nan_count);
                    if (flag) atomicAdd(&sm.stats[kStatSuccess], 1u);
                    if (!point_in_image(k2.x, k2.y, lv)) atomicAdd(&sm.stats[kStatOutOfImage], 1u);
                    state = ST_FETCH;
                }
            }
        }
    }

    __syncthreads();
    if (tid < kStatCount && sm.stats[tid]) atomicAdd(&args.stats[tid], (unsigned long long)sm.stats[tid]);
}

constexpr int kLaneThreads = 128;
constexpr int kLaneMinCtas = 2;

}  // namespace

bool lane_kernel_supports(const SolverArgs &args) {
    return args.patch_lo == LO && args.patch_hi == HI && !args.inverse && args.max_iters >= 1 && args.max_iters <= 15;
}

cudaError_t launch_klt_lane(const PyramidView &pyr, const SolverArgs &args, int sm_count, cudaStream_t stream) {
    if (args.n_total <= 0) return cudaSuccess;
    auto kernel = klt_lane_kernel<kLaneThreads, kLaneMinCtas>;
    const size_t smem = sizeof(LaneSmem<kLaneThreads>);
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int grid = sm_count * kLaneMinCtas;
    const int needed = (args.n_total + kLaneThreads - 1) / kLaneThreads;
    if (grid > needed) grid = needed;
    kernel<<<grid, kLaneThreads, smem, stream>>>(pyr, args);
    return cudaGetLastError();
}

}  // namespace legoklt
