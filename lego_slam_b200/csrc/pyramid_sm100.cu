// pyramid_sm100.cu -- K1: fused image-pyramid build for a batch of 8-bit images (sm_100a).
//
// Replaces the pyramid part of LKOpticalFlow4Layer, src/algorithm.cpp:140-154:
//     level i = cv::resize(level i-1, cv::Size(cols*0.5, rows*0.5))      // INTER_LINEAR, CV_8UC1
// OpenCV is third party (not under /root/reference); its 8-bit linear resize is integer fixed point
// and is restated here (pinned bit-exact against Python cv2 by the oracle tests):
//     f = (float)((d+0.5)*scale-0.5); s = floor(f); f -= s;   (x only: clamp s to [0,sn-1] with f=0)
//     a0 = rint((1-f)*2048), a1 = rint(f*2048)
//     h(x)  = S[sx]*a0 + S[sx+1]*a1                                        (int32, per source row)
//     dst   = (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2
//
// B200 design: HBM-bound on paper (level 0 read once, levels 1..L-1 written once: 619,601 B per 1241x376
// image at L=4), instruction-bound in practice, so the work is split where the bytes are:
//   pyramid_l01_kernel   level 0 -> 1 (three quarters of the bytes): streaming, a warp per output row, 16-byte
//                        loads, dp2a horizontal pass, no shared memory, no barrier; writes the level-0 aprons.
//   pyramid_band_kernel  levels 2.. from level 1 (L2-resident by then): a CTA owns a band of top-level rows,
//                        stages the level-1 rows it needs by ONE bulk copy (cp.async.bulk + mbarrier), derives
//                        every coarser level in shared memory and writes each band back by bulk stores; writes
//                        the aprons of levels 1...
// (The first version fused all levels in the band kernel: 1475 instructions per warp of which 370 were resize
// arithmetic -- per-level set-up, barriers and tiny phases dominated; profiles/README.md.)
// The coefficient tables are built once per shape on the host and live in global memory (L1/L2 hits).
#include <algorithm>
#include <cmath>
#include <vector>

#include "klt_kernels.h"

namespace legoklt {

bool pyramid_level_sizes(int cols, int rows, int levels, int *lcols, int *lrows) {
    lcols[0] = cols;
    lrows[0] = rows;
    for (int l = 1; l < levels; ++l) {
        lcols[l] = (int)(lcols[l - 1] * 0.5);  // cv::Size(int * double) truncates
        lrows[l] = (int)(lrows[l - 1] * 0.5);
        if (lcols[l] <= 0 || lrows[l] <= 0) return false;
    }
    return true;
}

namespace {

struct HostAxis {
    std::vector<int> ofs;
    std::vector<short2> coef;
};

HostAxis build_axis(int sn, int dn, bool horizontal) {
    HostAxis t;
    t.ofs.resize(dn);
    t.coef.resize(dn);
    const double inv_scale = (double)dn / sn;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (horizontal) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= sn - 1) { s = sn - 1; f = 0.f; }
        }
        t.ofs[d] = s;
        t.coef[d].x = (short)std::lrintf((1.f - f) * 2048.f);
        t.coef[d].y = (short)std::lrintf(f * 2048.f);
    }
    return t;
}

inline int clip_row(int v, int n) { return v < 0 ? 0 : (v < n ? v : n - 1); }

struct PyrKernelParams {
    ResizeTables tab[kMaxLevels];
    int smem_off[kMaxLevels];  // byte offset of level l's band (pixel 0 of its first row) in dynamic smem
    const int2 *bands;         // [n_bands][kMaxLevels] {first row, row count} of every level's band (host-built)
    int img0, nimg;            // image range of this launch (chunked batches)
};

__device__ __forceinline__ int d_clip(int v, int n) { return v < 0 ? 0 : (v < n ? v : n - 1); }

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (a0 * byte0 + a1 * byte1) of the low / high byte pair of `bytes`, a = {a0, a1} as two 16-bit fields.
__device__ __forceinline__ int dot2_lo(uint32_t a, uint32_t bytes) {
    int d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(bytes), "r"(0));
    return d;
}
__device__ __forceinline__ int dot2_hi(uint32_t a, uint32_t bytes) {
    int d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(bytes), "r"(0));
    return d;
}

// Vertical pass + rounding of cv::resize (see the file header) for one pixel.
template <bool Y2>
__device__ __forceinline__ uint32_t vertical(int h0, int h1, int b0, int b1) {
    if (Y2) return (uint32_t)(((h0 >> 10) + (h1 >> 10) + 2) >> 2);  // b0 = b1 = 1024: (1024*(h>>4))>>16 == h>>10
    return (uint32_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

// One destination row of level k: the lanes of the calling warp part stride over the 4-pixel groups.
//   X2 (horizontal scale exactly 2, weights 1024/1024) and Y2 (same vertically) are per-level properties.
//   X2 && Y2: (p00+p01+p10+p11+2)>>2 with two pixels per 32-bit word (SWAR on 16-bit lanes).
//   otherwise: a regular group (taps of pixel j start at sx0 + 2j: everywhere except where the tap offset steps
//   by one, and in the partial group at the row end) takes its 8 source bytes with two funnel shifts and does the
//   horizontal pass with dp2a; irregular groups take the per-pixel path.
template <bool X2, bool Y2>
__device__ __forceinline__ void resize_row(const uint8_t *__restrict__ s0, const uint8_t *__restrict__ s1,
                                           uint32_t *__restrict__ drow, const ResizeTables &tb, int b0, int b1,
                                           int dcols, int src_last, int g0, int gstep) {
    const int ng = (dcols + 3) >> 2;
    for (int g = g0; g < ng; g += gstep) {
        const int c0 = 4 * g;
        uint32_t packed;
        int info = X2 ? ((c0 + 3 < dcols) ? 2 * c0 : -1) : __ldg(tb.gofs + g);
        if (info >= 0) {
            uint32_t d00, d01, d10, d11;  // source bytes sx0..sx0+7 of the two source rows
            if (X2) {
                const uint2 t0 = *reinterpret_cast<const uint2 *>(s0 + info);
                const uint2 t1 = *reinterpret_cast<const uint2 *>(s1 + info);
                d00 = t0.x, d01 = t0.y, d10 = t1.x, d11 = t1.y;
            } else {
                const int base = info & ~3, sh = (info & 3) * 8;
                const uint32_t *q0 = reinterpret_cast<const uint32_t *>(s0 + base);
                const uint32_t *q1 = reinterpret_cast<const uint32_t *>(s1 + base);
                const uint32_t a0 = q0[0], a1 = q0[1], a2 = q0[2], c0w = q1[0], c1w = q1[1], c2w = q1[2];
                d00 = __funnelshift_r(a0, a1, sh), d01 = __funnelshift_r(a1, a2, sh);
                d10 = __funnelshift_r(c0w, c1w, sh), d11 = __funnelshift_r(c1w, c2w, sh);
            }
            if (X2 && Y2) {
                // even + odd bytes per 16-bit lane, both rows, +2, >>2: two output pixels per word
                const uint32_t sA = __byte_perm(d00, 0u, 0x4240) + __byte_perm(d00, 0u, 0x4341) +
                                    __byte_perm(d10, 0u, 0x4240) + __byte_perm(d10, 0u, 0x4341) + 0x00020002u;
                const uint32_t sB = __byte_perm(d01, 0u, 0x4240) + __byte_perm(d01, 0u, 0x4341) +
                                    __byte_perm(d11, 0u, 0x4240) + __byte_perm(d11, 0u, 0x4341) + 0x00020002u;
                packed = __byte_perm(sA >> 2, sB >> 2, 0x6420);
            } else {
                uint4 cf;
                if (X2) cf = make_uint4(0x04000400u, 0x04000400u, 0x04000400u, 0x04000400u);
                else cf = __ldg(reinterpret_cast<const uint4 *>(tb.xcoef) + g);
                const uint32_t v0 = vertical<Y2>(dot2_lo(cf.x, d00), dot2_lo(cf.x, d10), b0, b1);
                const uint32_t v1 = vertical<Y2>(dot2_hi(cf.y, d00), dot2_hi(cf.y, d10), b0, b1);
                const uint32_t v2 = vertical<Y2>(dot2_lo(cf.z, d01), dot2_lo(cf.z, d11), b0, b1);
                const uint32_t v3 = vertical<Y2>(dot2_hi(cf.w, d01), dot2_hi(cf.w, d11), b0, b1);
                packed = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
            }
            drow[g] = packed;
        }
    }
    // irregular groups (tap offset steps inside the group; partial group at the row end): one pixel per lane
    uint8_t *dbytes = reinterpret_cast<uint8_t *>(drow);
    for (int q = g0; q < 4 * tb.n_girr; q += gstep) {
        const int col = 4 * __ldg(tb.girr + (q >> 2)) + (q & 3);
        if (col < dcols) {
            const int sx = __ldg(tb.xofs + col);
            const int sx1 = min(sx + 1, src_last);
            const short2 cf = __ldg(tb.xcoef + col);
            const int h0 = (int)s0[sx] * cf.x + (int)s0[sx1] * cf.y;
            const int h1 = (int)s1[sx] * cf.x + (int)s1[sx1] * cf.y;
            dbytes[col] = (uint8_t)vertical<false>(h0, h1, b0, b1);
        }
    }
}

// Aprons of one row held in shared memory (LevelView: 32 bytes left of column 0 replicate it, columns > cols
// replicate column cols-1); the wrap byte at column `cols` is written by the caller once the row below exists.
__device__ __forceinline__ void smem_row_aprons(uint8_t *row, int cols, int pitch, int lane) {
    const uint32_t first = row[0], last = row[cols - 1];
    if (lane < kApronL / 4) reinterpret_cast<uint32_t *>(row - kApronL)[lane] = first * 0x01010101u;
    for (int c = cols + 1 + lane; c < pitch - kApronL; c += 32) row[c] = (uint8_t)last;
}

// Fused pyramid + aprons.  One CTA owns a band of top-level rows of one image:
//   1. the level-0 rows it needs arrive in shared memory by ONE bulk copy (cp.async.bulk, mbarrier completion):
//      a band is contiguous in the pitched layout, so staging costs no instructions;
//   2. level k is derived from level k-1 in shared memory, including its row aprons;
//   3. each level's band leaves by bulk stores (shared -> global), again contiguous.
// Bands of neighbouring CTAs may overlap by a row at the finer levels (both write identical bytes).  The only
// byte a band cannot know is the wrap byte of its LAST row (first pixel of the row below): it is excluded from
// the band's stores and written by the CTA whose band starts with that row.
__global__ void __launch_bounds__(256)
pyramid_band_kernel(const __grid_constant__ PyramidView pyr, const __grid_constant__ PyrKernelParams kp) {
    extern __shared__ __align__(128) uint8_t smem[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
    const int L = pyr.levels;
    const int set = blockIdx.y >= kp.nimg ? 1 : 0;
    const int img = kp.img0 + blockIdx.y - set * kp.nimg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kWarps = 8;
    // Row band {first row, row count} of every level for this CTA (built on the host: a band contains the rows the
    // next-coarser band reads AND reaches the first row of the next CTA's band, so that the bands of one level
    // tile it completely -- with truncated level sizes (135 -> 67) the coarser level does not read every finer row).
    const int2 *band = kp.bands + (size_t)blockIdx.x * kMaxLevels;
    const int2 band0 = __ldg(band);
    const int lo0 = band0.x, nr0 = band0.y;

    // ---- 1. level-0 band: bulk copy global -> shared
    const LevelView &l0 = pyr.lv[0];
    uint8_t *g0 = l0.base[set] + (size_t)img * l0.slot + (size_t)lo0 * l0.pitch;
    uint8_t *S0 = smem + kp.smem_off[0];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t total = (uint32_t)nr0 * (uint32_t)l0.pitch;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(total) : "memory");
        for (uint32_t off = 0; off < total; off += 16384u) {
            const uint32_t n = min(16384u, total - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(S0 + off)),
                         "l"(g0 + off), "r"(n), "r"(smem_addr(bar))
                         : "memory");
        }
    }
    __syncthreads();  // barrier initialised before anybody polls it
    {
        uint32_t ok = 0, spins = 0;
        while (!ok) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(ok)
                : "r"(smem_addr(bar)), "r"(0)
                : "memory");
            if (!ok && ++spins > (1u << 24)) __trap();  // a copy that never lands must not hang the GPU
        }
    }

    // ---- level-0 aprons: straight to global (level 0 itself is input; only the apron bytes are written)
    for (int r = warp; r < nr0; r += kWarps) {
        const int R = lo0 + r;
        const uint8_t *srow = S0 + (size_t)r * l0.pitch;
        uint8_t *grow = g0 + (size_t)r * l0.pitch;
        const uint32_t first = srow[0], last = srow[l0.cols - 1];
        uint32_t wrap;  // data[R*step + cols] of the reference's flat addressing (algorithm.h:48,53)
        if (l0.step != l0.cols) wrap = srow[l0.cols];  // inside the caller's row padding, uploaded with the row
        else if (R + 1 >= l0.rows) wrap = 0u;
        else wrap = (r + 1 < nr0) ? srow[l0.pitch] : __ldg(grow + l0.pitch);
        if (lane < kApronL / 4) reinterpret_cast<uint32_t *>(grow - kApronL)[lane] = first * 0x01010101u;
        for (int c = l0.cols + lane; c < l0.pitch - kApronL; c += 32) grow[c] = (uint8_t)(c == l0.cols ? wrap : last);
    }

    // ---- 2./3. coarser levels
    int lo_src = lo0;
#pragma unroll 1
    for (int k = 1; k < L; ++k) {
        const LevelView &ld = pyr.lv[k];
        const LevelView &ls = pyr.lv[k - 1];
        const uint8_t *S = smem + kp.smem_off[k - 1];
        uint8_t *D = smem + kp.smem_off[k];
        const ResizeTables &tb = kp.tab[k];
        const int2 bk = __ldg(band + k);
        const int lo = bk.x, nr = bk.y;
        const int dpitch = ld.pitch, dcols = ld.cols, drows = ld.rows, spitch = ls.pitch, srows = ls.rows;
        // few rows (coarse levels): 2^lg warps share a row
        int lg = 0;
        while ((2 << lg) * nr <= kWarps) ++lg;
        const int gfirst = (warp & ((1 << lg) - 1)) * 32 + lane, gstep = 32 << lg;
        const int mode = (tb.x_exact2 ? 2 : 0) | (tb.y_exact2 ? 1 : 0);
        for (int r = warp >> lg; r < nr; r += kWarps >> lg) {
            const int R = lo + r;
            const int yo = __ldg(tb.yofs + R);
            const short2 b = __ldg(tb.ycoef + R);
            const uint8_t *s0 = S + (d_clip(yo, srows) - lo_src) * spitch;
            const uint8_t *s1 = S + (d_clip(yo + 1, srows) - lo_src) * spitch;
            uint32_t *drow = reinterpret_cast<uint32_t *>(D + r * dpitch);
            if (mode == 3) resize_row<true, true>(s0, s1, drow, tb, b.x, b.y, dcols, ls.cols - 1, gfirst, gstep);
            else if (mode == 1) resize_row<false, true>(s0, s1, drow, tb, b.x, b.y, dcols, ls.cols - 1, gfirst, gstep);
            else if (mode == 2) resize_row<true, false>(s0, s1, drow, tb, b.x, b.y, dcols, ls.cols - 1, gfirst, gstep);
            else resize_row<false, false>(s0, s1, drow, tb, b.x, b.y, dcols, ls.cols - 1, gfirst, gstep);
        }
        __syncthreads();
        for (int r = warp; r < nr; r += kWarps) smem_row_aprons(D + r * dpitch, dcols, dpitch, lane);
        // wrap bytes (levels >= 1 are continuous, step == cols): first pixel of the row below, 0 after the last row
        uint8_t *gD = ld.base[set] + (size_t)img * ld.slot + (size_t)lo * dpitch;
        if (tid < nr) {
            if (lo + tid + 1 >= drows) D[tid * dpitch + dcols] = 0;
            else if (tid + 1 < nr) D[tid * dpitch + dcols] = D[(tid + 1) * dpitch];
        } else if (tid == 32 * (kWarps - 1) && lo > 0) {
            gD[dcols - dpitch] = D[0];  // wrap byte of the row above this band
        }
        // every thread's shared-memory writes must be visible to the async proxy that reads them for the bulk store
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // band out: [left apron of the first row, column c16 of the last row) in one piece, the rest of the last
        // row by plain stores that skip the wrap byte unless it is known (last row of the image)
        const int c16 = dcols & ~15;
        if (tid == 0) {
            const uint32_t total = (uint32_t)((nr - 1) * dpitch + c16 + kApronL);
            for (uint32_t off = 0; off < total; off += 16384u) {
                const uint32_t n = min(16384u, total - off);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gD - kApronL + off),
                             "r"(smem_addr(D - kApronL + off)), "r"(n)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else if (warp == 1) {
            const bool own_wrap = (lo + nr >= drows);
            const uint8_t *srow = D + (nr - 1) * dpitch;
            uint8_t *grow = gD + (size_t)(nr - 1) * dpitch;
            for (int c = c16 + lane; c < dpitch - kApronL; c += 32)
                if (c != dcols || own_wrap) grow[c] = srow[c];
        }
        lo_src = lo;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem is read until here
}

// ------------------------------------------------------------------------------------------------
// Level 0 -> level 1 (three quarters of all pyramid bytes) as a pure streaming kernel: no shared memory, no
// block barrier.  A warp owns one level-1 row; a lane takes 8 output pixels = 16 source bytes of each of the two
// source rows with one aligned 16-byte load per row.  A 0.5x level has taps sx = 2c + delta(c), delta in {0, 1}
// (dn = floor(sn/2), so the scale is 2 or 2 + 1/dn): where delta = 1 the 16 bytes start one byte later, and the
// missing byte comes from the next lane's load by shuffle.  The chunk in which delta steps (one per row when sn
// is odd) and the partial chunk at the row end take the per-pixel path.  The CTA also writes the level-0 row
// aprons of the 16 source rows under it (level-1 aprons are written by the band kernel that builds levels 2..).
// ------------------------------------------------------------------------------------------------
struct L01Params {
    const int *yofs;       // [rows1]
    const short2 *ycoef;   // [rows1]
    const int *cofs;       // [ceil(cols1/8)] first tap of a regular 8-pixel chunk (taps sx0 + 2j), else -1
    const int *irr;        // the irregular chunks (same for every row)
    int n_irr;
    const int *xofs;       // [cols1] (padded to 8)
    const short2 *xcoef;   // [cols1] (padded to 8)
    int img0, nimg;
};

// Output rows per warp: the lane keeps its chunk column and walks kL01Rows consecutive output rows with it, so that
// the chunk's coefficients and tap offset are loaded once, the 2 * kL01Rows 16-byte loads of a trip are all in flight
// together (the first version, one row per warp, was load-latency bound: DRAM 47 % of peak, issue slots 55 % active,
// 3 trips per warp) and the CTA count drops by the same factor.
#ifndef PYR_L01_ROWS
#define PYR_L01_ROWS 4
#endif
constexpr int kL01Rows = PYR_L01_ROWS;
constexpr int kL01RowsPerCta = 8 * kL01Rows;

template <bool X2Y2, bool Y2>
__global__ void __launch_bounds__(256)
pyramid_l01_kernel(const __grid_constant__ PyramidView pyr, const __grid_constant__ L01Params p) {
    const LevelView &l0 = pyr.lv[0];
    const LevelView &l1 = pyr.lv[1];
    const int set = blockIdx.y >= p.nimg ? 1 : 0;
    const int img = p.img0 + blockIdx.y - set * p.nimg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t *src = l0.base[set] + (size_t)img * l0.slot;
    uint8_t *dst = l1.base[set] + (size_t)img * l1.slot;
    const int spitch = l0.pitch, scols = l0.cols, srows = l0.rows;

    const int R0 = kL01RowsPerCta * blockIdx.x + kL01Rows * warp;
    if (R0 < l1.rows) {
    const int dcols = l1.cols, nchunks = (dcols + 7) >> 3, src_last = scols - 1;
    const int nrows = min(kL01Rows, l1.rows - R0);
    // The warp's work items are (row, chunk) pairs, row-major, a lane takes every 32nd: 78 chunks x 4 rows fill 10 trips
    // to 97 % (a lane per chunk column and a trip per row filled 81 %).  The chunk's coefficient loads hit L1.
    const int n_items = nrows * nchunks;
    for (int item = lane; item < n_items; item += 32) {
        int k = 0;
#pragma unroll
        for (int q = 1; q < kL01Rows; ++q) k += (item >= q * nchunks) ? 1 : 0;
        const int i = item - k * nchunks, R = R0 + k;
        const int yo = __ldg(p.yofs + R);
        const short2 b = __ldg(p.ycoef + R);
        const uint8_t *s0 = src + (size_t)d_clip(yo, srows) * spitch;
        const uint8_t *s1 = src + (size_t)d_clip(yo + 1, srows) * spitch;
        // chunk i = 16 source bytes per source row + the word after them (for delta = 1)
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(s0) + i);
        const uint4 c = __ldg(reinterpret_cast<const uint4 *>(s1) + i);
        const uint32_t na = __ldg(reinterpret_cast<const uint32_t *>(s0) + 4 * i + 4);  // (the row's allocation extends 48 bytes
        const uint32_t nc = __ldg(reinterpret_cast<const uint32_t *>(s1) + 4 * i + 4);  //  past the last pixel)
        const int info = __ldg(p.cofs + i);
        if (info >= 0) {
            const int sh = (info - 16 * i) * 8;  // 0 or 8 (host-checked)
            uint2 out;
            const uint32_t w0[4] = {__funnelshift_r(a.x, a.y, sh), __funnelshift_r(a.y, a.z, sh), __funnelshift_r(a.z, a.w, sh),
                                    __funnelshift_r(a.w, na, sh)};
            const uint32_t w1[4] = {__funnelshift_r(c.x, c.y, sh), __funnelshift_r(c.y, c.z, sh), __funnelshift_r(c.z, c.w, sh),
                                    __funnelshift_r(c.w, nc, sh)};
            if (X2Y2) {
                uint32_t t[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    t[q] = (__byte_perm(w0[q], 0u, 0x4240) + __byte_perm(w0[q], 0u, 0x4341) + __byte_perm(w1[q], 0u, 0x4240) +
                            __byte_perm(w1[q], 0u, 0x4341) + 0x00020002u) >> 2;
                out.x = __byte_perm(t[0], t[1], 0x6420);
                out.y = __byte_perm(t[2], t[3], 0x6420);
            } else {
                const uint4 cf0 = __ldg(reinterpret_cast<const uint4 *>(p.xcoef) + 2 * i);
                const uint4 cf1 = __ldg(reinterpret_cast<const uint4 *>(p.xcoef) + 2 * i + 1);
                const uint32_t cf[8] = {cf0.x, cf0.y, cf0.z, cf0.w, cf1.x, cf1.y, cf1.z, cf1.w};
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int h0 = (j & 1) ? dot2_hi(cf[j], w0[j >> 1]) : dot2_lo(cf[j], w0[j >> 1]);
                    const int h1 = (j & 1) ? dot2_hi(cf[j], w1[j >> 1]) : dot2_lo(cf[j], w1[j >> 1]);
                    v[j] = vertical<Y2>(h0, h1, b.x, b.y);
                }
                out.x = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
                out.y = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
            }
            reinterpret_cast<uint2 *>(dst + (size_t)R * l1.pitch)[i] = out;
        }
    }
    // The irregular chunks (where the tap offset steps, and the partial chunk at the row end): one pixel per lane, the
    // rows of the warp flattened the same way.
    const int n_irr_px = 8 * p.n_irr;
    for (int item = lane; item < nrows * n_irr_px; item += 32) {
        int k = 0;
#pragma unroll
        for (int q = 1; q < kL01Rows; ++q) k += (item >= q * n_irr_px) ? 1 : 0;
        const int q = item - k * n_irr_px, R = R0 + k;
        const int col = 8 * __ldg(p.irr + (q >> 3)) + (q & 7);
        if (col < dcols) {
            const int yo = __ldg(p.yofs + R);
            const short2 b = __ldg(p.ycoef + R);
            const uint8_t *s0 = src + (size_t)d_clip(yo, srows) * spitch;
            const uint8_t *s1 = src + (size_t)d_clip(yo + 1, srows) * spitch;
            const int sx = __ldg(p.xofs + col);
            const int sx1 = min(sx + 1, src_last);
            const short2 cf = __ldg(p.xcoef + col);
            const int h0 = (int)s0[sx] * cf.x + (int)s0[sx1] * cf.y;
            const int h1 = (int)s1[sx] * cf.x + (int)s1[sx1] * cf.y;
            dst[(size_t)R * l1.pitch + col] = (uint8_t)vertical<false>(h0, h1, b.x, b.y);
        }
    }
    }

    // ---- level-0 aprons of the 2 * kL01RowsPerCta source rows under this CTA: a lane per row, as few warps as that
    // takes (every warp doing two rows cost a quarter of the kernel's instructions, mostly 64-bit address
    // arithmetic: profiles/)
    constexpr int kApronWarps = (2 * kL01RowsPerCta + 15) / 16;   // 16 rows per warp: lanes 0-15 left, 16-31 right aprons
    static_assert(kApronWarps <= 8, "apron rows per CTA");
    if (warp >= 8 - kApronWarps) {
        const int Rr = 2 * kL01RowsPerCta * blockIdx.x + 16 * (warp - (8 - kApronWarps)) + (lane & 15);
        if (Rr < srows) {
            uint8_t *grow = const_cast<uint8_t *>(src) + (size_t)Rr * spitch;
            const uint32_t first = grow[0], last = grow[scols - 1];
            uint32_t wrap;  // data[R*step + cols] of the reference's flat addressing (algorithm.h:48,53)
            if (l0.step != scols) wrap = grow[scols];  // inside the caller's row padding, uploaded with the row
            else wrap = (Rr + 1 < srows) ? grow[spitch] : 0u;
            if (lane < 16) {  // left apron: 32 bytes, 16-byte aligned
                const uint32_t f4 = first * 0x01010101u;
                reinterpret_cast<uint4 *>(grow - kApronL)[0] = make_uint4(f4, f4, f4, f4);
                reinterpret_cast<uint4 *>(grow - kApronL)[1] = make_uint4(f4, f4, f4, f4);
            } else {          // right apron: wrap byte, then the last pixel replicated up to the next row's left apron
                const int end = spitch - kApronL;
                int c = scols;
                grow[c++] = (uint8_t)wrap;
                for (; (c & 3) && c < end; ++c) grow[c] = (uint8_t)last;
                const uint32_t l4 = last * 0x01010101u;
                for (; c + 4 <= end; c += 4) *reinterpret_cast<uint32_t *>(grow + c) = l4;
                for (; c < end; ++c) grow[c] = (uint8_t)last;
            }
        }
    }
}

// Row aprons: one WARP per row; grid = (row blocks of all levels, image * set) so that no thread divides by a run-time
// size (the first version spent 180 instructions per row on index arithmetic: profiles/README.md).  32 + ~50
// bytes per row, written by the lanes in parallel; tiny next to the pyramid itself, and it lets the solver
// stage windows with unconditional aligned loads.
__global__ void __launch_bounds__(256) apron_kernel(const __grid_constant__ PyramidView pyr, int img0, int nimg) {
    int level = 0, rb = blockIdx.x;  // blockIdx.x runs over the row blocks of all levels, level 0 first
    for (; level < pyr.levels - 1; ++level) {
        const int nb = (pyr.lv[level].rows + 7) >> 3;
        if (rb < nb) break;
        rb -= nb;
    }
    const LevelView &lv = pyr.lv[level];
    const int row = rb * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= lv.rows) return;
    const int set = blockIdx.y >= nimg ? 1 : 0;
    const int k = img0 + blockIdx.y - set * nimg;
    uint8_t *rp = lv.base[set] + (size_t)k * lv.slot + (size_t)row * lv.pitch;
    const uint32_t first = rp[0], last = rp[lv.cols - 1];
    // data[row*step + cols] of the reference's flat addressing
    uint32_t wrap;
    if (lv.step == lv.cols)
        wrap = (row + 1 < lv.rows) ? rp[lv.pitch] : 0u;
    else
        wrap = rp[lv.cols];  // inside the caller's row padding, which was uploaded with the row
    if (lane < kApronL / 4) reinterpret_cast<uint32_t *>(rp - kApronL)[lane] = first * 0x01010101u;
    for (int c = lv.cols + lane; c < lv.pitch - kApronL; c += 32) rp[c] = (uint8_t)(c == lv.cols ? wrap : last);
}

// Ingest: level 0 arrives from the host as tight rows (`step` bytes each, arbitrary alignment -- KITTI's
// 1241 is odd); a strided cudaMemcpy2D into the pitched layout ran at ~9 GB/s, a plain copy runs at PCIe
// speed, so rows are re-pitched on the device.  One warp per row chunk: aligned 4-byte loads, funnel
// shift by the row's misalignment, aligned 4-byte stores.
__global__ void __launch_bounds__(256)
ingest_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst_base, int step, int pitch, long long row0,
              long long n_rows, int rows_per_image, unsigned long long slot) {
    const int words = (step + 3) >> 2;
    const long long total = n_rows * words;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long lrow = i / words;
        const int w = (int)(i - lrow * words);
        const long long row = row0 + lrow;                         // row index within the whole image set
        const size_t s = (size_t)row * step + 4 * (size_t)w;      // byte offset of this word in the tight buffer
        const size_t sa = s & ~(size_t)3;
        const int sh = (int)(s & 3) * 8;
        const uint32_t lo = *reinterpret_cast<const uint32_t *>(src + sa);
        const uint32_t hi = sh ? *reinterpret_cast<const uint32_t *>(src + sa + 4) : 0u;
        const long long img = row / rows_per_image;
        const int r = (int)(row - img * rows_per_image);
        *reinterpret_cast<uint32_t *>(dst_base + (size_t)img * slot + (size_t)r * pitch + 4 * (size_t)w) =
            __funnelshift_r(lo, hi, sh);
    }
}

// Dataset::NextFrame's cv::resize(img, out, cv::Size(), 0.5, 0.5, cv::INTER_NEAREST) (src/dataset.cpp:75-77; OpenCV
// is third party: dsize = cvRound(size * 0.5), source index = min(floor(d * 2), size - 1); pinned against
// Python cv2 by tests): full-resolution tight rows -> the pitched level 0 of a device image.  One thread per 4
// output pixels.
__global__ void __launch_bounds__(256)
half_nearest_kernel(const uint8_t *__restrict__ src, int scols, int srows, size_t sstep, uint8_t *__restrict__ dst,
                    int dcols, int drows, int dpitch) {
    const int groups = (dcols + 3) >> 2;
    const long long total = (long long)groups * drows;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / groups), g = (int)(i - (long long)y * groups);
        const uint8_t *srow = src + (size_t)min(2 * y, srows - 1) * sstep;
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = 4 * g + j;
            if (x < dcols) packed |= (uint32_t)__ldg(srow + min(2 * x, scols - 1)) << (8 * j);
        }
        *reinterpret_cast<uint32_t *>(dst + (size_t)y * dpitch + 4 * g) = packed;  // (spill-over lands in the row's apron)
    }
}

}  // namespace

cudaError_t launch_half_nearest(const uint8_t *full, int full_cols, int full_rows, size_t full_step, const LevelView &l0,
                                cudaStream_t stream) {
    const long long total = (long long)((l0.cols + 3) / 4) * l0.rows;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    half_nearest_kernel<<<grid, 256, 0, stream>>>(full, full_cols, full_rows, full_step, l0.base[0], l0.cols, l0.rows, l0.pitch);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_ingest(const uint8_t *tight, const LevelView &l0, int set, int img0, int n_images,
                          cudaStream_t stream) {  // (single-image handles pass set = 0)
    const long long n_rows = (long long)n_images * l0.rows;
    if (n_rows <= 0) return cudaSuccess;
    const long long total = n_rows * ((l0.step + 3) >> 2);
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    ingest_kernel<<<grid, 256, 0, stream>>>(tight, l0.base[set], l0.step, l0.pitch, (long long)img0 * l0.rows, n_rows,
                                            l0.rows, l0.slot);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_aprons(const PyramidView &pyr, int img0, int nimg, cudaStream_t stream, int n_sets) {
    if (nimg <= 0 || pyr.levels <= 0) return cudaSuccess;
    int row_blocks = 0;
    for (int l = 0; l < pyr.levels; ++l) row_blocks += (pyr.lv[l].rows + 7) / 8;
    dim3 grid(row_blocks, n_sets * nimg);
    apron_kernel<<<grid, 256, 0, stream>>>(pyr, img0, nimg);
    note_launch();
    return cudaGetLastError();
}

cudaError_t pyramid_plan_create(int cols, int rows, int levels, const int *pitch, PyramidPlan *plan,
                                bool band_kernel_only) {
    *plan = PyramidPlan();
    plan->levels = levels;
    if (!pyramid_level_sizes(cols, rows, levels, plan->cols, plan->rows)) return cudaErrorInvalidValue;
    for (int l = 0; l < levels; ++l) plan->pitch[l] = pitch[l];
    if (levels == 1) return cudaSuccess;

    // ---- tables (one blob; every table padded to a multiple of 4 entries, 16-byte aligned) ----
    std::vector<HostAxis> hx(levels), hy(levels);
    size_t blob = 0;
    auto pad4 = [](int n) { return (size_t)((n + 3) & ~3); };
    auto pad8 = [](int n) { return (size_t)((n + 7) & ~7); };  // x tables: the streaming kernel reads 8 entries at a time
    for (int l = 1; l < levels; ++l) {
        hx[l] = build_axis(plan->cols[l - 1], plan->cols[l], true);
        hy[l] = build_axis(plan->rows[l - 1], plan->rows[l], false);
        blob += pad8(plan->cols[l]) * 8 + pad4(plan->rows[l]) * 8 + 2 * pad4((plan->cols[l] + 3) / 4) * 4 +
                2 * pad4((plan->cols[l] + 7) / 8) * 4;
    }
    std::vector<uint8_t> host(blob, 0);
    cudaError_t err = cudaMalloc(&plan->table_blob, blob);
    if (err != cudaSuccess) return err;
    size_t off = 0;
    for (int l = 1; l < levels; ++l) {
        uint8_t *dbase = static_cast<uint8_t *>(plan->table_blob);
        ResizeTables &t = plan->tab[l];
        const int nc = plan->cols[l], nr = plan->rows[l];
        t.xofs = reinterpret_cast<int *>(dbase + off);
        memcpy(&host[off], hx[l].ofs.data(), 4 * (size_t)nc);
        off += pad8(nc) * 4;
        t.xcoef = reinterpret_cast<short2 *>(dbase + off);
        memcpy(&host[off], hx[l].coef.data(), 4 * (size_t)nc);
        off += pad8(nc) * 4;
        t.yofs = reinterpret_cast<int *>(dbase + off);
        memcpy(&host[off], hy[l].ofs.data(), 4 * (size_t)nr);
        off += pad4(nr) * 4;
        t.ycoef = reinterpret_cast<short2 *>(dbase + off);
        memcpy(&host[off], hy[l].coef.data(), 4 * (size_t)nr);
        off += pad4(nr) * 4;
        t.x_exact2 = 1;
        for (int c = 0; c < nc; ++c)
            if (hx[l].ofs[c] != 2 * c || hx[l].coef[c].x != 1024 || hx[l].coef[c].y != 1024) t.x_exact2 = 0;
        t.y_exact2 = 1;
        for (int r = 0; r < nr; ++r)
            if (hy[l].ofs[r] != 2 * r || hy[l].coef[r].x != 1024 || hy[l].coef[r].y != 1024) t.y_exact2 = 0;
        // per 4-pixel group: first tap of the group if its pixels tap sx0, sx0+2, sx0+4, sx0+6 (the kernel's
        // regular path), else -1 (tap offset steps inside the group, or partial group at the row end)
        const int ng = (nc + 3) / 4;
        std::vector<int> gofs(pad4(ng), -1);
        for (int g = 0; g < ng; ++g) {
            const int c0 = 4 * g;
            bool regular = c0 + 3 < nc && hx[l].ofs[c0] >= 0;
            for (int j = 1; regular && j < 4; ++j) regular = hx[l].ofs[c0 + j] == hx[l].ofs[c0] + 2 * j;
            if (regular) gofs[g] = hx[l].ofs[c0];
        }
        t.gofs = reinterpret_cast<int *>(dbase + off);
        memcpy(&host[off], gofs.data(), 4 * (size_t)ng);
        off += pad4(ng) * 4;
        // per 8-pixel chunk (level 0 -> 1 streaming kernel): first tap if the chunk is complete, its pixels tap
        // sx0 + 2j and sx0 is 0..1 bytes past the chunk's aligned 16 source bytes
        const int nch = (nc + 7) / 8;
        std::vector<int> cofs(pad4(nch), -1);
        for (int i = 0; i < nch; ++i) {
            const int c0 = 8 * i;
            bool regular = c0 + 7 < nc && hx[l].ofs[c0] >= 16 * i && hx[l].ofs[c0] <= 16 * i + 1;
            for (int j = 1; regular && j < 8; ++j) regular = hx[l].ofs[c0 + j] == hx[l].ofs[c0] + 2 * j;
            if (regular) cofs[i] = hx[l].ofs[c0];
        }
        t.cofs = reinterpret_cast<int *>(dbase + off);
        memcpy(&host[off], cofs.data(), 4 * (size_t)nch);
        off += pad4(nch) * 4;
        // the irregular groups / chunks as lists (the kernels give them one pixel per lane after the regular ones)
        std::vector<int> girr, irr;
        for (int g = 0; g < ng; ++g)
            if (gofs[g] < 0) girr.push_back(g);
        for (int i = 0; i < nch; ++i)
            if (cofs[i] < 0) irr.push_back(i);
        t.girr = reinterpret_cast<int *>(dbase + off);
        t.n_girr = (int)girr.size();
        memcpy(&host[off], girr.data(), 4 * girr.size());
        off += pad4(ng) * 4;
        t.irr = reinterpret_cast<int *>(dbase + off);
        t.n_irr = (int)irr.size();
        memcpy(&host[off], irr.data(), 4 * irr.size());
        off += pad4(nch) * 4;
    }
    err = cudaMemcpy(plan->table_blob, host.data(), blob, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) return err;

    if (!band_kernel_only) {
        // level 0 -> 1: streaming kernel (tables of level 1 above); levels 2..: band kernel on the pyramid seen from level 1
        plan->sub = new PyramidPlan();
        return pyramid_plan_create(plan->cols[1], plan->rows[1], levels - 1, pitch + 1, plan->sub, true);
    }

    // ---- bands: the largest top-row band whose staging fits the shared-memory budget; every band's row range
    // at every level goes into a device table (the kernel used to derive it with a chain of dependent loads) ----
    const int top = levels - 1;
    const size_t budget = 48 * 1024;  // measured on B200 (1241x376, L=4): 40-48 KB bands best, 64-100 KB 5 % slower
    std::vector<int2> bands;
    for (int tr = 32; tr >= 1; tr >>= 1) {
        int maxr[kMaxLevels] = {0};
        bands.clear();
        for (int t0 = 0; t0 < plan->rows[top]; t0 += tr) {
            const bool last_band = t0 + tr >= plan->rows[top];
            int lo = t0, hi = std::min(t0 + tr, plan->rows[top]), next_lo = t0 + tr;
            int2 rec[kMaxLevels] = {};
            rec[top] = make_int2(lo, hi - lo);
            maxr[top] = std::max(maxr[top], hi - lo);
            for (int k = top; k >= 1; --k) {
                int nlo = clip_row(hy[k].ofs[lo], plan->rows[k - 1]);
                int needed_hi = clip_row(hy[k].ofs[hi - 1] + 1, plan->rows[k - 1]) + 1;
                next_lo = last_band ? plan->rows[k - 1] : clip_row(hy[k].ofs[next_lo], plan->rows[k - 1]);
                lo = nlo;
                hi = std::max(needed_hi, next_lo);
                rec[k - 1] = make_int2(lo, hi - lo);
                maxr[k - 1] = std::max(maxr[k - 1], hi - lo);
            }
            bands.insert(bands.end(), rec, rec + kMaxLevels);
        }
        size_t bytes = 128;  // mbarrier
        for (int l = 0; l < levels; ++l) {
            bytes = ((bytes + 15) & ~(size_t)15) + kApronL;  // the first row's left apron precedes the band
            plan->smem_off[l] = (int)bytes;
            plan->max_rows[l] = maxr[l];
            bytes += (size_t)maxr[l] * plan->pitch[l];
        }
        plan->smem_bytes = bytes;
        plan->top_rows_per_cta = tr;
        if (bytes <= budget || tr == 1) break;
    }
    plan->n_bands = (int)(bands.size() / kMaxLevels);
    err = cudaMalloc(&plan->band_tab, bands.size() * sizeof(int2));
    if (err != cudaSuccess) return err;
    err = cudaMemcpy(plan->band_tab, bands.data(), bands.size() * sizeof(int2), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) return err;
    if (plan->smem_bytes > 220 * 1024) return cudaErrorInvalidValue;
    return cudaSuccess;
}

void pyramid_plan_destroy(PyramidPlan *plan) {
    if (plan->sub) {
        pyramid_plan_destroy(plan->sub);
        delete plan->sub;
    }
    if (plan->table_blob) cudaFree(plan->table_blob);
    if (plan->band_tab) cudaFree(plan->band_tab);
    *plan = PyramidPlan();
}

namespace {

// ------------------------------------------------------------------------------------------------
// Levels 2 and 3 from level 1 when both halvings are EXACT (source = 2 x destination in both directions: every
// (a + b + c + d + 2) >> 2 of cv::resize's fixed-point arithmetic) -- the reference's 1241x376 chain (620x188 -> 310x94
// -> 155x47).  Streaming, no shared memory, no barrier: a work item is 8 level-1 pixels of four consecutive rows
// (4 aligned 8-byte loads) -> 4 + 4 level-2 pixels (two rows) -> 2 level-3 pixels, all in one lane's registers, two
// pixels per 32-bit operation; a warp takes kX2Groups row groups and its lanes every 32nd (group, chunk) item.  Then
// the warp writes the row aprons of the 7 rows of each of its groups (4 of level 1, 2 of level 2, 1 of level 3), one
// row per lane: first / last pixel and the wrap byte (first pixel of the next row) are recomputed from level 1, so no
// value has to travel between lanes, warps or CTAs.  Replaces the band kernel for this shape (44 -> 31 us per 512
// images, 36.7 M -> 13.3 M warp instructions: the band kernel spends 63 instructions per output pixel on staging,
// barriers and band bookkeeping).
// ------------------------------------------------------------------------------------------------
#ifndef PYR_X2_GROUPS
#define PYR_X2_GROUPS 1         // row groups per warp (measured 1 / 2 / 4: pyramid 0.113 / 0.116 / 0.122 ms per 512 images)
#endif
#ifndef PYR_X2_WARPS
#define PYR_X2_WARPS 8
#endif
constexpr int kX2Groups = PYR_X2_GROUPS, kX2Warps = PYR_X2_WARPS;
static_assert(7 * kX2Groups <= 32, "one apron row per lane");

// two output pixels (16-bit lanes) from a word of 4 source pixels of each of two rows
__device__ __forceinline__ uint32_t avg2x2_pairs(uint32_t a, uint32_t b) {
    return (__byte_perm(a, 0u, 0x4240) + __byte_perm(a, 0u, 0x4341) + __byte_perm(b, 0u, 0x4240) + __byte_perm(b, 0u, 0x4341) +
            0x00020002u) >> 2;
}
// one level-2 pixel from level 1 (x even: two aligned 16-bit loads)
__device__ __forceinline__ uint32_t avg2x2_at(const uint8_t *row, int pitch, int x) {
    const uint32_t a = *reinterpret_cast<const uint16_t *>(row + x), b = *reinterpret_cast<const uint16_t *>(row + pitch + x);
    return ((a & 0xffu) + (a >> 8) + (b & 0xffu) + (b >> 8) + 2u) >> 2;
}
// one level-3 pixel from level 1 through its four level-2 pixels (x a multiple of 4: four aligned 32-bit loads)
__device__ __forceinline__ uint32_t avg4x4_at(const uint8_t *row, int pitch, int x) {
    const uint32_t a = *reinterpret_cast<const uint32_t *>(row + x), b = *reinterpret_cast<const uint32_t *>(row + pitch + x);
    const uint32_t c = *reinterpret_cast<const uint32_t *>(row + 2 * (size_t)pitch + x);
    const uint32_t d = *reinterpret_cast<const uint32_t *>(row + 3 * (size_t)pitch + x);
    const uint32_t u0 = avg2x2_pairs(a, b), u1 = avg2x2_pairs(c, d);   // two level-2 pixels each, in the low bytes of 16-bit lanes
    return ((u0 & 0xffu) + ((u0 >> 16) & 0xffu) + (u1 & 0xffu) + ((u1 >> 16) & 0xffu) + 2u) >> 2;
}

__global__ void __launch_bounds__(32 * kX2Warps)
pyramid_x2x2_kernel(const __grid_constant__ PyramidView pyr, int img0, int nimg) {
    const LevelView &l1 = pyr.lv[0], &l2 = pyr.lv[1], &l3 = pyr.lv[2];   // (the view starts at level 1)
    const int set = blockIdx.y >= nimg ? 1 : 0;
    const int img = img0 + blockIdx.y - set * nimg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *g1 = l1.base[set] + (size_t)img * l1.slot;
    uint8_t *g2 = l2.base[set] + (size_t)img * l2.slot;
    uint8_t *g3 = l3.base[set] + (size_t)img * l3.slot;
    const int p1 = l1.pitch, p2 = l2.pitch, p3 = l3.pitch;
    const int G0 = (blockIdx.x * kX2Warps + warp) * kX2Groups;   // first row group (= level-3 row) of this warp
    if (G0 >= l3.rows) return;
    const int ngroups = min(kX2Groups, l3.rows - G0);
    const int nchunks = (l1.cols + 7) >> 3;
    for (int item = lane; item < ngroups * nchunks; item += 32) {
        int k = 0;
#pragma unroll
        for (int q = 1; q < kX2Groups; ++q) k += (item >= q * nchunks) ? 1 : 0;
        const int j = item - k * nchunks, g = G0 + k;
        const uint8_t *r = g1 + (size_t)(4 * g) * p1 + 8 * j;
        const uint2 a = *reinterpret_cast<const uint2 *>(r), b = *reinterpret_cast<const uint2 *>(r + p1);
        const uint2 c = *reinterpret_cast<const uint2 *>(r + 2 * p1), d = *reinterpret_cast<const uint2 *>(r + 3 * p1);
        const uint32_t u0 = __byte_perm(avg2x2_pairs(a.x, b.x), avg2x2_pairs(a.y, b.y), 0x6420);   // level-2 row 2g, 4 pixels
        const uint32_t u1 = __byte_perm(avg2x2_pairs(c.x, d.x), avg2x2_pairs(c.y, d.y), 0x6420);   // level-2 row 2g + 1
        const uint32_t v = __byte_perm(avg2x2_pairs(u0, u1), 0u, 0x4420);                          // level-3 row g, 2 pixels
        uint8_t *o2 = g2 + (size_t)(2 * g) * p2 + 4 * j, *o3 = g3 + (size_t)g * p3 + 2 * j;
        if (4 * j + 4 <= l2.cols) {
            *reinterpret_cast<uint32_t *>(o2) = u0;
            *reinterpret_cast<uint32_t *>(o2 + p2) = u1;
            *reinterpret_cast<uint16_t *>(o3) = (uint16_t)v;
        } else {   // the half chunk at the end of a row (level-3 width odd): only the pixels that exist, byte by byte
            for (int q = 0; 4 * j + q < l2.cols; ++q) {
                o2[q] = (uint8_t)(u0 >> (8 * q));
                o2[p2 + q] = (uint8_t)(u1 >> (8 * q));
            }
            for (int q = 0; 2 * j + q < l3.cols; ++q) o3[q] = (uint8_t)(v >> (8 * q));
        }
    }
    // ---- aprons: lane = one row of one group (rows 0..3: level 1, 4..5: level 2, 6: level 3)
    if (lane < 7 * ngroups) {
        const int k = lane / 7, q = lane - 7 * k, g = G0 + k;
        uint32_t first, last, wrap;
        uint8_t *grow;
        int cols, pitch;
        if (q < 4) {
            const int y = 4 * g + q;
            const uint8_t *row = g1 + (size_t)y * p1;
            cols = l1.cols, pitch = p1, grow = g1 + (size_t)y * p1;
            first = row[0], last = row[cols - 1];
            wrap = (y + 1 < l1.rows) ? row[p1] : 0u;
        } else if (q < 6) {
            const int y = 2 * g + (q - 4);
            const uint8_t *row = g1 + (size_t)(2 * y) * p1;
            cols = l2.cols, pitch = p2, grow = g2 + (size_t)y * p2;
            first = avg2x2_at(row, p1, 0), last = avg2x2_at(row, p1, 2 * (cols - 1));
            wrap = (y + 1 < l2.rows) ? avg2x2_at(row + 2 * (size_t)p1, p1, 0) : 0u;
        } else {
            auto px3 = [&](int y, int x) -> uint32_t { return avg4x4_at(g1 + (size_t)(4 * y) * p1, p1, 4 * x); };
            cols = l3.cols, pitch = p3, grow = g3 + (size_t)g * p3;
            first = px3(g, 0), last = px3(g, cols - 1);
            wrap = (g + 1 < l3.rows) ? px3(g + 1, 0) : 0u;
        }
        const uint32_t f4 = first * 0x01010101u, l4 = last * 0x01010101u;
        reinterpret_cast<uint4 *>(grow - kApronL)[0] = make_uint4(f4, f4, f4, f4);   // left apron: 32 bytes, 16-byte aligned
        reinterpret_cast<uint4 *>(grow - kApronL)[1] = make_uint4(f4, f4, f4, f4);
        const int end = pitch - kApronL;   // right apron: wrap byte, then the last pixel up to the next row's left apron
        int cc = cols;
        grow[cc++] = (uint8_t)wrap;
        for (; (cc & 3) && cc < end; ++cc) grow[cc] = (uint8_t)last;
        for (; cc + 4 <= end; cc += 4) *reinterpret_cast<uint32_t *>(grow + cc) = l4;
        for (; cc < end; ++cc) grow[cc] = (uint8_t)last;
    }
}

bool x2x2_applies(const PyramidPlan &sub) {   // `sub`: the plan of levels 1.. (its level 0 is level 1)
    return sub.levels == 3 && sub.tab[1].x_exact2 && sub.tab[1].y_exact2 && sub.tab[2].x_exact2 && sub.tab[2].y_exact2 &&
           sub.cols[0] == 4 * sub.cols[2] && sub.rows[0] == 4 * sub.rows[2];
}

cudaError_t launch_x2x2_kernel(const PyramidPlan &sub, const PyramidView &up, int img0, int nimg, cudaStream_t stream, int n_sets) {
    const int groups_per_cta = kX2Groups * kX2Warps;
    dim3 grid((sub.rows[2] + groups_per_cta - 1) / groups_per_cta, n_sets * nimg);
    pyramid_x2x2_kernel<<<grid, 32 * kX2Warps, 0, stream>>>(up, img0, nimg);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_band_kernel(const PyramidPlan &plan, const PyramidView &pyr, int img0, int nimg, cudaStream_t stream,
                               int n_sets) {
    if (plan.levels <= 1) return launch_aprons(pyr, img0, nimg, stream, n_sets);  // no levels to build: aprons only
    PyrKernelParams kp;
    for (int l = 0; l < kMaxLevels; ++l) {
        kp.tab[l] = plan.tab[l];
        kp.smem_off[l] = plan.smem_off[l];
    }
    kp.bands = static_cast<const int2 *>(plan.band_tab);
    kp.img0 = img0;
    kp.nimg = nimg;
    if (plan.smem_bytes > 48 * 1024) {
        // the attribute belongs to the current device's context, not to the calling thread: set it on every launch
        // (a host-side table lookup) instead of caching it per thread
        cudaError_t err = cudaFuncSetAttribute(pyramid_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)plan.smem_bytes);
        if (err != cudaSuccess) return err;
    }
    dim3 grid(plan.n_bands, n_sets * nimg);
    pyramid_band_kernel<<<grid, 256, plan.smem_bytes, stream>>>(pyr, kp);
    note_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_pyramid(const PyramidPlan &plan, const PyramidView &pyr, int img0, int nimg, cudaStream_t stream,
                           int n_sets) {
    if (nimg <= 0) return cudaSuccess;
    if (!plan.sub) return launch_band_kernel(plan, pyr, img0, nimg, stream, n_sets);
    // level 0 -> 1 (+ level-0 aprons)
    L01Params p;
    const ResizeTables &t = plan.tab[1];
    p.yofs = t.yofs;
    p.ycoef = t.ycoef;
    p.cofs = t.cofs;
    p.xofs = t.xofs;
    p.xcoef = t.xcoef;
    p.irr = t.irr;
    p.n_irr = t.n_irr;
    p.img0 = img0;
    p.nimg = nimg;
    // kL01RowsPerCta level-1 rows and the twice as many level-0 rows under them per CTA; one more CTA if an odd last
    // level-0 row is left over
    const int ctas = std::max((plan.rows[1] + kL01RowsPerCta - 1) / kL01RowsPerCta,
                              (plan.rows[0] + 2 * kL01RowsPerCta - 1) / (2 * kL01RowsPerCta));
    dim3 grid(ctas, n_sets * nimg);
    if (t.x_exact2 && t.y_exact2) pyramid_l01_kernel<true, true><<<grid, 256, 0, stream>>>(pyr, p);
    else if (t.y_exact2) pyramid_l01_kernel<false, true><<<grid, 256, 0, stream>>>(pyr, p);
    else pyramid_l01_kernel<false, false><<<grid, 256, 0, stream>>>(pyr, p);
    note_launch();
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    // levels 2..: the same pyramid seen from level 1 (the band kernel also writes the aprons of its level 0)
    PyramidView up;
    up.levels = pyr.levels - 1;
    up.n_images = pyr.n_images;
    for (int l = 0; l < up.levels; ++l) up.lv[l] = pyr.lv[l + 1];
    if (x2x2_applies(*plan.sub)) return launch_x2x2_kernel(*plan.sub, up, img0, nimg, stream, n_sets);
    return launch_band_kernel(*plan.sub, up, img0, nimg, stream, n_sets);
}

}  // namespace legoklt
