// klt_solver_lane.cu -- LEGO_KLT_KERNEL_LANE: one THREAD per feature, persistent per-thread state
// machine, all pyramid levels fused (sm_100a).  The patch bounds and the mode are compile-time and this source is
// compiled once per configuration: here forward mode (the reference's call sites pass inverse = false:
// src/frontend_g2o.cpp:473,515), 7x7 = -3..3 (src/algorithm.cpp:40 half_patch_size = 3); in klt_solver_lane_p8.cu 8x8
// = -4..3, in klt_solver_lane_p11.cu 11x11 = -5..5, in klt_solver_lane_inv.cu the reference's inverse mode (7x7; see
// LANE_INVERSE below).  The numbers quoted below are those of the forward 7x7 instance.
//
// Replaces LKOpticalFlow1Layer + LKOpticalFlowTracker::calcLKOpticalFlow (src/algorithm.cpp:11-125)
// and the coarse-to-fine loop of LKOpticalFlow4Layer (:158-205).  Two kernels:
//
//   klt_template_kernel  one thread per (feature, level): the template patch I1 of
//                        src/algorithm.cpp:65 (GetPixelValue(img1, kx+x, ky+y)), constant over the
//                        Gauss-Newton passes of a level, written once as P*P floats + a regularity flag
//                        (level-major records, transposed through shared memory for coalesced stores).
//   klt_lane_kernel      one thread per feature: all passes of all levels.
//
// Why not warp-per-feature (measured, profiles/r01_*): that mapping spends ~850 warp instructions per
// Gauss-Newton pass -- shuffles for six fp64 reductions, a serial 2x2 solve on every lane, per-sample
// index arithmetic -- for 49 pixels of useful work.  Here every lane runs a whole pass for its own
// feature (no shuffles, no idle lanes in the solve).  Divergence in iteration counts (1..10 per level)
// is removed by a state machine: each trip of the main loop is exactly one pass for every runnable
// thread; a thread that converges moves to its next level / next feature while its neighbours keep
// iterating.  Level set-up (template fetch + img2 window staging) is a per-thread, divergent section run
// whenever at least LANE_BATCH threads of the warp need it (measured best: 1).
//
// Shared memory, per thread, in 16-byte granules  [granule][T]:  img2 window 12 rows x 32 bytes (16-byte aligned
// origin; the border semantics come from the row aprons) and the template, 7 rows x 8 floats -- 608 bytes per thread
// (the row factors of a pass are formed inside its row loop, not parked in shared memory), five CTAs of 64 threads per
// SM.  A level set-up is 38 16-byte cp.async copies from global memory.
//
// What bounds it (profiles/README.md, round 2): throughput, not latency -- issue slots are 57 % active whatever the
// occupancy, most of its instruction kinds cost ~2 issue cycles per sub-partition on this part, so time follows the
// instruction count: ~2,200 per warp-pass, of which ~1,240 the row loop.
//
// Bit-fidelity contract (same as the warp kernel): every fp32 value entering the sums is bit-identical
// to the reference's; only the ORDER of the fp64 additions differs (row-major here, x-outer there; all
// products are exact in fp64).  The (P+2)^2 sample grid is shared between "centre of pixel x+1" and
// "+1 tap of pixel x" only when that is provably what the reference computes:
//   (a) the pixel offsets c of one sub-pass share the rounding error e_c of float(kx+c) (axis_families):
//       normally e_c = 0 for all c and a pass is one trip; near a power of two the patch splits into two
//       families per axis, each with its own grid, and the pass takes one trip per family combination with
//       the other pixels masked out (sub-pixel source keypoints, i.e. tracked points fed back by the
//       reference's TrackLastFrame, hit this on ~10 % of their passes; such features are listed by the template
//       kernel and tracked by the FAM = true instance of the kernel);
//   (b) the double coordinate is not within 16 ulp64 of an fp32 rounding midpoint (checked per pass,
//       per grid column/row), so the <=2 ulp64 differences between the reference's three ways of
//       forming a coordinate cannot change the rounded float -- or it is EXACTLY on a midpoint and all
//       double sums are provably exact (TwoSum), which is the common first-pass case (kx + float dx).
// The border semantics of algorithm.h:42-55 ARE part of the fast path: clamped coordinates get the
// factors (1,0) over replicated border pixels; a sample in the last-pixel sliver x in (cols-1, cols)
// reads the reference's flat-address neighbour (first pixel of the next row); both come from the row
// aprons every device image carries (LevelView), so window staging never branches on the border;
// y in (rows-1, rows) reads zeros below the image (the oracle's definition of the reference's
// out-of-buffer read).  A feature that violates (a) or (b), or whose window leaves the 32-pixel apron,
// is appended to a deferred list and finished by the exact warp kernel; its partial work is dropped.
#include <type_traits>

#include "klt_kernels.h"

namespace legoklt {

namespace {

// Patch bounds are compile-time: this file is compiled once per supported patch (klt_solver_lane_p8.cu and
// klt_solver_lane_p11.cu include it with other LANE_PATCH_LO/HI, window / CTA shapes and a symbol suffix).
#ifndef LANE_PATCH_LO
#define LANE_PATCH_LO (-3)
#define LANE_PATCH_HI 3
#endif
#ifndef LANE_SUFFIX
#define LANE_SUFFIX
#endif
#define LANE_CAT2(a, b) a##b
#define LANE_CAT(a, b) LANE_CAT2(a, b)
#define LANE_FN(name) LANE_CAT(name, LANE_SUFFIX)
constexpr int LO = LANE_PATCH_LO, HI = LANE_PATCH_HI, P = HI - LO + 1, G = P + 2;
static_assert(P >= 3 && P <= 12, "family masks hold 12 bits per axis");
// LANE_INVERSE: the reference's inverse mode INCLUDING its stale Jacobian (src/algorithm.cpp:57,74-80,83):
// the Jacobian of a pixel is the img1 gradient at the template position, refreshed only in the first pass of a level;
// every later pass uses the Jacobian of the LAST pixel (x = y = HI) of that first pass for all pixels, and H is the
// first pass's.  The template kernel then also stores the two img1 gradient differences per pixel (three planes per
// record), the solver samples img2 at the P x P pixel centres only (no +-1 taps), and after the first pass of a level
// a thread overwrites its gradient planes with the last pixel's values, so that every pass runs the same code.
#ifndef LANE_INVERSE
#define LANE_INVERSE 0
#endif
constexpr bool kInverse = LANE_INVERSE != 0;
constexpr int kG0 = kInverse ? 1 : 0;                          // first grid index the solver samples on img2
constexpr int kNS = kInverse ? P : G;                          // img2 samples per grid row / sampled grid rows
constexpr int kNP = (kNS + 1) / 2;                             // sample pairs per grid row: samples (2j, 2j+1), j < kNP
constexpr int kRowWords = (2 * kNP + 1 + 3) / 4 + 1;           // window words a row step reads: 2 kNP + 1 pixels at a
                                                              // byte offset of 0..3
constexpr unsigned kPMask = (1u << P) - 1u;                    // one bit per patch column / row
constexpr int kFamY = 12, kFamSub = 24;                        // family word: x mask | y mask << kFamY | sub-pass << kFamSub
constexpr unsigned kPMask2 = kPMask | (kPMask << kFamY);
#ifndef LANE_WROWS
#define LANE_WROWS (G + 3)
#endif
#ifndef LANE_WWORDS
#define LANE_WWORDS 8
#endif
#ifndef LANE_QUEUE
#define LANE_QUEUE 64          // (measured 32 / 64: 1.465 / 1.44 ms per 512,000 features)
#endif
#ifndef LANE_WUNIT
#define LANE_WUNIT 4           // words per window copy unit: 16 / 8 / 4-byte cp.async copies, origin aligned to the unit
#endif
// img2 window per thread: LANE_WROWS rows x LANE_WWORDS words, origin aligned to one copy unit.
constexpr int kWin2Rows = LANE_WROWS, kWin2Words = LANE_WWORDS, kWinUnit = LANE_WUNIT;
static_assert((kWinUnit == 1 || kWinUnit == 2 || kWinUnit == 4) && kWin2Words % kWinUnit == 0, "window copy unit");
constexpr int kWinAlign = 4 * kWinUnit;
constexpr int kWinSlackL = (kWinUnit == 4) ? 2 : 3;             // footprint starts kWinSlackL..kWinSlackL+kWinAlign-1 bytes in
constexpr int kWinSpare = LANE_WROWS - (G + 1);                // window rows beyond the footprint
constexpr int kWinSlackT = kWinSpare / 2;                      // rows above the footprint (an odd spare row: see the stager)
static_assert(kWin2Rows >= G + 1 && kWin2Words * 4 >= G + 1 + kWinSlackL + kWinAlign - 1,
              "the window must hold the footprint at every alignment (a footprint that never fits would re-stage forever)");
constexpr int kTplRows = kInverse ? G + 1 : P + 1;            // img1 window rows in the template kernel
// The per-thread shared memory is made of 16-byte granules, [granule][T], so that a level set-up is 38 16-byte
// cp.async copies straight from global memory, all in flight together, no register staging (the first layout,
// [word][T] with conflict-free 32-bit accesses, needed ~170 set-up instructions incl. 121 32-bit stores: 7 % slower
// overall, profiles/README.md).  Template rows are padded to whole granules and read with 128-bit loads; window
// words are read one by one (4-way bank conflicts, affordable: the shared-memory pipe is far from saturated).
constexpr int kTplPitch = (P + 3) / 4 * 4;                    // floats per template row, in the record and in shared memory
constexpr int kI1Count = P * kTplPitch;                       // template floats, index y*kTplPitch + x
constexpr int kQueue = LANE_QUEUE;                            // feature ring entries per warp (power of two, >= 32)
static_assert(kQueue >= 32 && (kQueue & (kQueue - 1)) == 0, "ring size");
constexpr int kRecPlanes = kInverse ? 3 : 1;                  // record planes: I1 [, img1 x difference, img1 y difference]
constexpr int kRecCount = kRecPlanes * kI1Count;
constexpr int kTplStride = ((kRecCount + 1 + 3) / 4 * 4) | 4;  // floats per (feature, level): planes + flag + pad; an ODD
                                                              // number of float4 (conflict-free transposed stores)
static_assert(kTplStride >= kRecCount + 1 && (kTplStride / 4) % 2 == 1 && kTplStride % 4 == 0, "template record stride");
#ifndef LANE_BATCH
#define LANE_BATCH 1
#endif
#ifndef LANE_T
#define LANE_T 64
#endif
#ifndef LANE_PINGPONG
#define LANE_PINGPONG 1
#endif
#ifndef LANE_PARK_SMEM
#define LANE_PARK_SMEM 1       // partial sums of two-family passes wait in shared memory (48 bytes per thread), not in registers
#endif
#ifndef LANE_FAM_CTAS
#define LANE_FAM_CTAS LANE_CTAS  // CTAs per SM of the family instance's grid (measured 1 / 2 / 3 / 4 / 5 per SM on 512,000
#endif                           // sub-pixel keypoints: 2.76 / 2.18 / 1.98 / 1.93 / 1.92 ms; CTAs without work exit at once)
#ifndef LANE_CTAS
#define LANE_CTAS 5
#endif
constexpr int kSetupBatch = LANE_BATCH;                       // blocked threads per warp that trigger set-up

enum : int { ST_FETCH = 0, ST_LEVEL = 1, ST_RUN = 2, ST_DONE = 3 };

template <int T>
struct LaneSmem {
    // One thread's window: copy units of kWinUnit words, [unit][T] (kWin2Words / kWinUnit per row); its record: 16-byte
    // granules, [granule][T] (kTplPitch / 4 per template row).
    static constexpr int kUnitsPerRow = kWin2Words / kWinUnit, kWinUnits = kWin2Rows * kUnitsPerRow;
    uint32_t win[kWinUnits][T][kWinUnit];
    uint4 rec[kRecCount / 4][T];
    // per-warp ring of fetched features: one global atomic + coalesced keypoint loads per 32 features
    float2 q_k1[T / 32][kQueue], q_k2[T / 32][kQueue];
    int q_id[T / 32][kQueue];
    unsigned stats[kStatCount];
#if LANE_PARK_SMEM
    double parked[6][T];  // partial sums between the sub-passes of a two-family pass
#endif
};

#ifndef LANE_ODD_MOV
#define LANE_ODD_MOV 0
#endif
#ifndef LANE_WIDEN_ALU
#define LANE_WIDEN_ALU 0       // experiment: how many of a pixel's three fp32 -> fp64 conversions avoid the XU pipe (F2F)
#endif
// fp32 -> fp64 by integer widening (exact; normal numbers on the ALU pipe, zero / denormal / inf / nan by the convert)
__device__ __forceinline__ double widen_alu(float f) {
    const uint32_t u = __float_as_uint(f);
    if (((u << 1) - 0x01000000u) >= 0xfe000000u) return (double)f;
    const int hi = (int)((((int)u >> 3) & 0x8fffffff) + 0x38000000);
    return __hiloint2double(hi, (int)(u << 29));
}

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

// One grid axis (columns or rows) of the shared sample grid.  S = kd + d, kd = double(float(k+c)) - c for the
// pixel offsets c of one coordinate family (see axis_families); the reference forms
// float(double(float(k+c)) + d [+-1]), which for that family is S + g - 4 up to 2 ulp64.
// Per grid index g, axis_factor() yields the two bilinear factors the reference would use:
//     om = 1 - frac   (weight of the tap at the integer coordinate)
//     fr = frac       (weight of the tap at integer coordinate + 1)
// with the nominal integer coordinate of index g being origin + g (consecutive), and the border
// semantics of algorithm.h:42-55 folded into the factors:
//   * X <  0      : clamped to 0        -> (1, 0); the window holds replicated border pixels
//   * X >= limit  : clamped to limit-1  -> (1, 0), or (0, 1) when origin+g == limit exactly, because
//                   window column `limit` holds the flat-address "wrap" pixel (see the stager) and the
//                   replicated border pixel sits one further
//   * sliver X in (limit-1, limit), not clamped: the +1 tap is data[.. + 1] on the next row (columns:
//     the wrap pixel, true factors) or past the buffer = 0 (rows: fr := 0, om = 1 - frac).
// Condition (b) of the header comment is tracked as a running minimum `near` over the indices of the axis (see
// axis_near): the pass is provably bit-identical on this grid iff axis_ok(near) afterwards.
// (The nominal coordinate needs no check: for |S| < 1e6 every D_g = fl64(S + n) lies within half an ulp64 of the
// real sum, so fl32(D_g) is in [origin + g, origin + g + 1] -- integers are fp32-representable -- and a factor pair
// (0, 1) at the nominal column equals the reference's (1, 0) one column further, term by term.)
struct Axis {
    double S;         // kd + d
    int origin;       // nominal integer coordinate of grid index 0: floor(S + LO - 1)
    float forigin;    // (float)origin
    unsigned tie;     // 1 if an EXACT rounding tie is harmless on this axis (see axis_begin), else 0
    bool in_range;    // |S| < 1e6
};

__device__ __forceinline__ Axis axis_begin(double kd, double d) {
    Axis ax;
    ax.S = kd + d;
    ax.in_range = fabs(ax.S) < 1.0e6;
    // An EXACT tie (D on an fp32 rounding midpoint) is safe when every double sum involved is exact:
    // then all of the reference's ways of forming the coordinate give the same double, and round-half-
    // even gives the same float.  TwoSum error of S, and 4 spare low bits so that S + c stays exact.
    const double bb = ax.S - kd;
    const double err = (kd - (ax.S - bb)) + (d - bb);
    ax.tie = ((err == 0.0) && ((__double2loint(ax.S) & 0xF) == 0) && (fabs(ax.S) >= 1.0)) ? 1u : 0u;
    ax.origin = __double2int_rd(ax.S + (double)(LO - 1));
    ax.forigin = (float)ax.origin;
    return ax;
}

// Distance of D's discarded mantissa bits from the fp32 rounding midpoint, folded into a running minimum:
// |dist| for a non-tie, and for an exact tie 0 (tie not provably harmless) or UINT_MAX (harmless).
__device__ __forceinline__ unsigned axis_near(unsigned near, double D, unsigned tie) {
    const int dist = (__double2loint(D) & 0x1FFFFFFF) - 0x10000000;
    return min(near, (unsigned)abs(dist) - tie);
}

// (b) holds on the axis iff no index came within 16 ulp64 of a midpoint (harmless ties excepted).
__device__ __forceinline__ bool axis_ok(unsigned near, unsigned tie) { return near >= 17u - tie; }

// n = (double)(LO - 1 + g), fg = (float)(origin + g); flimit = (float)limit, flast = (float)(limit - 1).
template <bool IS_ROW>
__device__ __forceinline__ void axis_factor(const Axis &ax, double n, float fg, float flimit, float flast, unsigned &near,
                                            float &om, float &fr) {
    const double D = ax.S + n;
    const float X = (float)D;
    near = axis_near(near, D, ax.tie);
    const float f_raw = __fadd_rn(X, -fg);
    const float o_raw = __fadd_rn(1.f, -f_raw);
    const bool clamped = (X < 0.f) || (X >= flimit);
    if (IS_ROW) {
        fr = ((X < 0.f) || (X > flast)) ? 0.f : f_raw;  // clamped, or taps below the last row read zeros
        om = clamped ? 1.f : o_raw;
    } else {
        const float fc = ((X >= flimit) && (fg == flimit)) ? 1.f : 0.f;  // the wrap column
        fr = clamped ? fc : f_raw;
        om = clamped ? __fadd_rn(1.f, -fc) : o_raw;
    }
}

// The whole axis at once (template kernel; columns of the solver).  Returns false if the pass cannot be proven
// bit-identical on this grid.
template <bool IS_ROW>
__device__ __forceinline__ bool grid_axis(double kd, double d, int limit, int &origin, float (&fr)[G], float (&om)[G]) {
    const Axis ax = axis_begin(kd, d);
    origin = ax.origin;
    const float flimit = (float)limit, flast = (float)(limit - 1);
    unsigned near = 0xFFFFFFFFu;
#pragma unroll
    for (int g = 0; g < G; ++g)
        axis_factor<IS_ROW>(ax, (double)(LO - 1 + g), __fadd_rn(ax.forigin, (float)g), flimit, flast, near, om[g], fr[g]);
    return ax.in_range && axis_ok(near, ax.tie);
}

// Coordinate families of one axis (condition (a), generalised).  The reference forms the sample coordinate of
// pixel offset c from float(k + c) (src/algorithm.cpp:65-73); near a power of two that float sum is rounded,
// i.e. float(k + c) = k + c + e_c with e_c != 0.  Offsets with the same e_c share one sample grid.  Family A
// (e = 0) always contains c = 0; maskB marks the offsets (bit c - LO) of family B, which share eps.
// Returns false if more than one non-zero e_c occurs (several powers of two inside the patch: left to the
// warp kernel).  Fast exit: if both ends are exact, every offset is (the end of larger magnitude decides).
__device__ __forceinline__ bool axis_families(float k, unsigned &maskB, double &eps) {
    const double kd = (double)k;
    maskB = 0u;
    eps = 0.0;
    if ((double)(k + (float)LO) == kd + (double)LO && (double)(k + (float)HI) == kd + (double)HI) return true;
    bool ok = true;
#pragma unroll
    for (int c = LO; c <= HI; ++c) {
        const double e = (double)(k + (float)c) - (kd + (double)c);
        if (e != 0.0) {
            if (maskB != 0u && e != eps) ok = false;
            eps = e;
            maskB |= 1u << (c - LO);
        }
    }
    return ok;
}

// A thread's window is kWin2Rows x 32 bytes of img2 with nominal origin (wx0 16-aligned, wy0).  Rows are clamped to
// the image; columns outside [0, cols) come from the row aprons (LevelView), so every copy is an aligned 16-byte
// copy -- provided the window lies inside the aprons:
__device__ __forceinline__ bool window_in_apron(const LevelView &lv, int wx0, int width_bytes = kWin2Words * 4) {
    return wx0 >= -kApronL && wx0 + width_bytes <= lv.pitch - kApronL;
}

// ---- packed FP32x2 variant of the row load / sample row (sm_100 FMUL2 / FADD2: two IEEE-rounded fp32 operations
// per instruction, bit-identical to the scalar ones).  A pixel row is kept as five even pairs (0,1)(2,3)..(8,9)
// and five odd pairs (1,2)(3,4)..(9,10), so that both taps of two adjacent samples are register pairs.
struct Row2 {
    float2 e[kNP], o[kNP];
};

// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (seen with the __fmul2_rn/__fadd2_rn intrinsics and with
// explicit .rn PTX, -fmad=false notwithstanding: 12 % of the features lost bit-identity, like any FMA contraction,
// SURVEY.md F9).  A sum whose addend is a product is therefore written as fma(product, 1.0, acc): the product is
// already rounded, product * 1.0 is exact, and an FMA cannot absorb a second multiplication.
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
    float2 c;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(rc));
    return c;
}

__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
    float2 c;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(rc));
    return c;
}

// acc + prod with one rounding, where prod is an already rounded product (see above).  `one` is 1.0f read from the
// kernel arguments: with a literal 1.0 ptxas simplifies fma(mul(a,b), 1, acc) to fma(a, b, acc), i.e. contracts.
__device__ __forceinline__ float2 add2_product(float2 acc, float2 prod, float one) {
    unsigned long long ra, rp, rc, one2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(one2) : "f"(one));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rp) : "f"(prod.x), "f"(prod.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rc) : "l"(rp), "l"(one2), "l"(ra));
    float2 c;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(rc));
    return c;
}

__device__ __forceinline__ float2 bytes_to_float2(uint32_t wa, int ka, uint32_t wb, int kb) {
    float2 p;
    p.x = __uint_as_float(__byte_perm(wa, 0x4B000000u, 0x7440u + ka));
    p.y = __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7440u + kb));
    return add2_rn(p, make_float2(-8388608.0f, -8388608.0f));
}

// Where a thread finds the window words of its footprint: row i, k-th word = base[i * row_stride + off[k]]
// (off[k]: the word's granule and its position inside it).
struct WinRef {
    const uint32_t *base;
    int row_stride;
    int off[kRowWords];
};

__device__ __forceinline__ void load_row10_packed(const WinRef &wr, int i, int sh, Row2 &row) {
    const uint32_t *p = wr.base + i * wr.row_stride;
    uint32_t w[kRowWords], b[kRowWords - 1];
#pragma unroll
    for (int k = 0; k < kRowWords; ++k) w[k] = p[wr.off[k]];
#pragma unroll
    for (int k = 0; k < kRowWords - 1; ++k) b[k] = __funnelshift_r(w[k], w[k + 1], sh);
#pragma unroll
    for (int j = 0; j < kNP; ++j) row.e[j] = bytes_to_float2(b[(2 * j) >> 2], (2 * j) & 3, b[(2 * j + 1) >> 2], (2 * j + 1) & 3);
#if LANE_ODD_MOV
    // experiment: the odd pairs (2j+1, 2j+2) re-use the even pairs' converted pixels (two moves instead of two PRMT + one
    // packed add); the last one needs pixel 2 kNP, converted alone
#pragma unroll
    for (int j = 0; j + 1 < kNP; ++j) row.o[j] = make_float2(row.e[j].y, row.e[j + 1].x);
    row.o[kNP - 1] = make_float2(row.e[kNP - 1].y, byte_to_float(b[(2 * kNP) >> 2], (2 * kNP) & 3));
#else
#pragma unroll
    for (int j = 0; j < kNP; ++j) row.o[j] = bytes_to_float2(b[(2 * j + 1) >> 2], (2 * j + 1) & 3, b[(2 * j + 2) >> 2], (2 * j + 2) & 3);
#endif
}

// Samples g = 2j, 2j+1 of one grid row (algorithm.h:51-56 per sample, products and sums individually rounded).
__device__ __forceinline__ void sample_row_packed(const float2 (&OMX)[kNP], const float2 (&XX)[kNP], float omy_r, float yy_r,
                                                  const Row2 &A, const Row2 &B, float one, float2 (&out)[kNP]) {
    const float2 omy2 = make_float2(omy_r, omy_r), yy2 = make_float2(yy_r, yy_r);
#pragma unroll
    for (int j = 0; j < kNP; ++j) {
        float2 r = mul2_rn(mul2_rn(OMX[j], omy2), A.e[j]);
        r = add2_product(r, mul2_rn(mul2_rn(XX[j], omy2), A.o[j]), one);
        r = add2_product(r, mul2_rn(mul2_rn(OMX[j], yy2), B.e[j]), one);
        r = add2_product(r, mul2_rn(mul2_rn(XX[j], yy2), B.o[j]), one);
        out[j] = r;
    }
}

__device__ __forceinline__ float pick(const float2 (&S)[kNP], int g) { return (g & 1) ? S[g >> 1].y : S[g >> 1].x; }

__device__ __noinline__ float sample_flat_cold(const uint8_t *img, const LevelView &lv, float x, float y) {
    return sample_flat(img, lv, x, y);
}

// Exact per-pixel pass in the reference formulation (src/algorithm.cpp:63-88: 5 GetPixelValue per pixel,
// flat addressing), for the rare pass whose shared sample grid cannot be proven bit-identical.  Cold code:
// kept out of line so that the hot loop stays inside the instruction cache.
__device__ __noinline__ void exact_pass(const uint8_t *img2, const LevelView &lv, const float *i1p, int ws, float kx,
                                        float ky, double dx, double dy, double (&sums)[6]) {
    double sb0 = 0, sb1 = 0, sc = 0, s00 = 0, s01 = 0, s11 = 0;
#pragma unroll 1
    for (int p = 0; p < P * P; ++p) {
        const int y = p / P, x = p - y * P;
        const float fx = kx + (float)(LO + x), fy = ky + (float)(LO + y);
        const double cx = (double)fx + dx, cy = (double)fy + dy;
        const int idx = y * kTplPitch + x;  // (ws: floats between consecutive words / granules of this thread)
        const float i1v = i1p[(idx >> 2) * ws + (idx & 3)];
        const double e = (double)(i1v - sample_flat_cold(img2, lv, (float)cx, (float)cy));
        const double gx = (double)(sample_flat_cold(img2, lv, (float)(cx + 1), (float)cy) -
                                   sample_flat_cold(img2, lv, (float)(cx - 1), (float)cy));
        const double gy = (double)(sample_flat_cold(img2, lv, (float)cx, (float)(cy + 1)) -
                                   sample_flat_cold(img2, lv, (float)cx, (float)(cy - 1)));
        sb0 = fma(e, gx, sb0);
        sb1 = fma(e, gy, sb1);
        sc = fma(e, e, sc);
        s00 = fma(gx, gx, s00);
        s01 = fma(gx, gy, s01);
        s11 = fma(gy, gy, s11);
    }
    sums[0] = sb0;
    sums[1] = sb1;
    sums[2] = sc;
    sums[3] = s00;
    sums[4] = s01;
    sums[5] = s11;
}

// The same for the inverse mode: only the img2 centre samples are taken per pixel in the reference formulation
// (src/algorithm.cpp:65-66); the gradient differences are the record's (first pass: img1 gradients, later: stale).
__device__ __noinline__ void exact_pass_inverse(const uint8_t *img2, const LevelView &lv, const float *rec, int ws, float kx,
                                                float ky, double dx, double dy, double (&sums)[6]) {
    double sb0 = 0, sb1 = 0, sc = 0, s00 = 0, s01 = 0, s11 = 0;
#pragma unroll 1
    for (int p = 0; p < P * P; ++p) {
        const int y = p / P, x = p - y * P;
        const float fx = kx + (float)(LO + x), fy = ky + (float)(LO + y);
        const double cx = (double)fx + dx, cy = (double)fy + dy;
        const int idx = y * kTplPitch + x;
        const float i1v = rec[(idx >> 2) * ws + (idx & 3)];
        const double e = (double)(i1v - sample_flat_cold(img2, lv, (float)cx, (float)cy));
        const double gx = (double)rec[((kI1Count + idx) >> 2) * ws + ((kI1Count + idx) & 3)];
        const double gy = (double)rec[((2 * kI1Count + idx) >> 2) * ws + ((2 * kI1Count + idx) & 3)];
        sb0 = fma(e, gx, sb0);
        sb1 = fma(e, gy, sb1);
        sc = fma(e, e, sc);
        s00 = fma(gx, gx, s00);
        s01 = fma(gx, gy, s01);
        s11 = fma(gy, gy, s11);
    }
    sums[0] = sb0;
    sums[1] = sb1;
    sums[2] = sc;
    sums[3] = s00;
    sums[4] = s01;
    sums[5] = s11;
}

// Level-l coordinate of a level-0 keypoint coordinate (src/algorithm.cpp:160-169 then :194 repeatedly):
// float(k * 2^-(L-1)), then exact doublings.
__device__ __forceinline__ float level_coord(float k0, int L, int level) {
    float k = (float)(k0 * (1.0 / (double)(1 << (L - 1))));
    for (int l = L - 1; l > level; --l) k = (float)((double)k / 0.5);
    return k;
}

// ------------------------------------------------------------------------------------------------
// Template kernel: I1 patch of every (feature, level).  Item i -> level = L-1 - i / n_total (coarse
// levels first), feature = i % n_total, so a warp works on neighbouring features of one level.
// ------------------------------------------------------------------------------------------------
// kTplPx = P + 2 consecutive pixels starting at byte `ox` (2..17) of a 32-byte row held in two uint4 registers: the
// word offset ox>>2 (0..4) is applied with a 3-stage select network (registers cannot be indexed
// dynamically), the byte offset with funnel shifts.  No shared memory: the first version of this kernel
// bounced every row through shared memory and was MIO-throttled (profiles/README.md).
constexpr int kTplPx = kInverse ? G + 1 : P + 2;   // (inverse mode: the whole G x G grid needs G + 1 pixels per row)
constexpr int kTplWords = (kTplPx + 3) / 4 + 1;   // words holding kTplPx pixels at a byte offset of 0..3
__device__ __forceinline__ void row_from_regs(const uint4 &q0, const uint4 &q1, int k, int sh, float (&row)[G + 1]) {
    const uint32_t w[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, 0u};
    const bool k4 = (k & 4) != 0, k2 = (k & 2) != 0, k1 = (k & 1) != 0;
    uint32_t t[kTplWords + 1], o[kTplWords], b[kTplWords - 1];
#pragma unroll
    for (int j = 0; j < kTplWords + 1; ++j) {
        const uint32_t a = w[min(j, 8)], b2 = w[min(j + 2, 8)], b4 = w[min(j + 4, 8)];
        t[j] = k4 ? b4 : (k2 ? b2 : a);
    }
#pragma unroll
    for (int j = 0; j < kTplWords; ++j) o[j] = k1 ? t[j + 1] : t[j];
#pragma unroll
    for (int j = 0; j < kTplWords - 1; ++j) b[j] = __funnelshift_r(o[j], o[j + 1], sh);
#pragma unroll
    for (int i = 0; i < kTplPx; ++i) row[i] = byte_to_float(b[i >> 2], i & 3);
}

template <int T>
__global__ void __launch_bounds__(T)
klt_template_kernel(const __grid_constant__ PyramidView pyr, const __grid_constant__ SolverArgs args) {
    // Output staging: a warp's 32 records (consecutive features of one level: 32 x 208 contiguous bytes in the
    // level-major layout) are transposed through shared memory and leave as 13 fully coalesced 512-byte stores.
    // Per-thread 16-byte stores at a 208-byte stride cost as much as all of the kernel's arithmetic (measured:
    // 316 us -> 154 us without the stores; without the LOADS still 313 us).  Record stride = 13 float4 (odd):
    // conflict-free both ways.
    __shared__ float4 stage[T / 32][32 * (kTplStride / 4)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = pyr.levels;
    // item space padded so that every level starts at a multiple of 32: a warp never straddles two levels
    const int n_pad = (args.n_total + 31) & ~31;
    const long long item = (long long)blockIdx.x * T + tid;
    const int level = L - 1 - (int)(item / n_pad);   // (coarse levels first)
    const int idx = (int)(item % n_pad);
    const bool in_range = level >= 0 && idx < args.n_total;
    const int f = args.f0 + (in_range ? idx : 0);
    const int img = f / args.n_per_pair;
    const bool valid = in_range && !slot_unused(args, f, img);  // (unused slots of a ragged batch get an empty record)
    const LevelView &lv = pyr.lv[valid ? level : 0];
    const float2 k0 = args.kp1[f];
    const float kx = level_coord(k0.x, L, level), ky = level_coord(k0.y, L, level);

#ifndef LANE_TPL_EARLY
#define LANE_TPL_EARLY 0
#endif
    // Experiment (measured, not adopted): request the img1 window rows BEFORE the grid factors are computed -- 43 % of
    // the kernel's stall samples wait for these loads right after issuing them.  1: L2 prefetches first (solver 1.55 ms
    // against 1.48); 2: the loads themselves first, held in registers across the factor computations (167 registers,
    // three CTAs per SM instead of four: 1.48, no gain -- the scattered 16-byte requests, not their latency, bound it).
    uint4 q[kTplRows][2];
    bool early = false;
    if (!kInverse && LANE_TPL_EARLY && valid && fabs((double)kx) < 1.0e6 && fabs((double)ky) < 1.0e6) {
        const int ixe = __double2int_rd((double)kx + (double)(LO - 1)), iye = __double2int_rd((double)ky + (double)(LO - 1));
        const int wx0 = (ixe - 2) & ~15, wy0 = iye + 1;
        early = window_in_apron(lv, wx0, 32);
        if (early) {
            const uint8_t *img1 = lv.base[0] + (size_t)img * lv.slot;
#pragma unroll
            for (int i = 0; i < kTplRows; ++i) {
                const int ry = min(max(wy0 + i, 0), lv.rows - 1);
                const uint4 *rp = reinterpret_cast<const uint4 *>(img1 + (ptrdiff_t)ry * lv.pitch + wx0);
                if (LANE_TPL_EARLY == 1) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 1));
                } else {
                    q[i][0] = __ldg(rp);
                    q[i][1] = __ldg(rp + 1);
                }
            }
        }
    }
    float xx[G], omx[G], yy[G], omy[G];
    int ixn = 0, iyn = 0;
    unsigned mBx, mBy;
    double ex, ey;
    bool regular = valid && axis_families(kx, mBx, ex) && axis_families(ky, mBy, ey);
    if (regular) {
        // float(kx + c) == (double)kx + c + e_c exactly: the template grid is the d = 0 grid of each family
        const bool okx = grid_axis<false>((double)kx, 0.0, lv.cols, ixn, xx, omx);
        regular = grid_axis<true>((double)ky, 0.0, lv.rows, iyn, yy, omy) && okx;
    }
    if (regular && mBx) {  // columns of family B take their factors from the grid shifted by eps
        float xb[G], ob[G];
        int ib = 0;
        regular = grid_axis<false>((double)kx + ex, 0.0, lv.cols, ib, xb, ob) && ib == ixn;
#pragma unroll
        for (int x = 0; x < P; ++x)
            if ((mBx >> x) & 1u) {
                xx[x + 1] = xb[x + 1];
                omx[x + 1] = ob[x + 1];
            }
    }
    if (regular && mBy) {
        float yb[G], ob[G];
        int ib = 0;
        regular = grid_axis<true>((double)ky + ey, 0.0, lv.rows, ib, yb, ob) && ib == iyn;
#pragma unroll
        for (int y = 0; y < P; ++y)
            if ((mBy >> y) & 1u) {
                yy[y + 1] = yb[y + 1];
                omy[y + 1] = ob[y + 1];
            }
    }
    // Patches up to 8x8 collect the record in registers and store it as float4; larger ones (121 values) go to the
    // staging memory value by value instead, or the kernel would not fit the register file.
    constexpr bool kDirect = (kRecCount > 64);
    float *mine_f = reinterpret_cast<float *>(&stage[warp][lane * (kTplStride / 4)]);
    float buf[kDirect ? 4 : kTplStride];
#pragma unroll
    for (int i = 0; i < (kDirect ? 4 : kTplStride); ++i) buf[i] = 0.f;
    if (regular) regular = window_in_apron(lv, (ixn - 2) & ~15, 32);
    if (kInverse && regular && (mBx | mBy)) regular = false;  // (the inverse instance has no family handling: exact warp kernel)
    // Class of the feature, decided by its levels with an atomicMax on its flag (values of earlier runs are smaller):
    //   4*epoch + 2  irregular on some level -> the exact warp kernel (deferred list)
    //   4*epoch + 1  two coordinate families on some level -> the lane kernel's SECOND phase (family list): those
    //                features need the masked row loop, and a warp runs it whenever one of its threads does, so they
    //                are tracked together, after the others
    //   smaller      the lane kernel's first phase.
    // A feature is appended to a list by the thread that raises its flag to that list's value; the lane kernel
    // re-checks the flag (a family feature may still be raised to the deferred class by another level).
    if (!valid) {
    } else if (!regular) {
        const int tag = 4 * args.epoch + 2;
        if (atomicMax(&args.feat_flag[f], tag) < tag) {
            args.defer_list[atomicAdd(args.defer_count, 1)] = f;
            atomicAdd(&args.stats[kStatDeferred], 1ull);
            atomicAdd(&args.stats[kStatDeferInexact], 1ull);
        }
    } else if (mBx | mBy) {
        const int tag = 4 * args.epoch + 1;
        if (atomicMax(&args.feat_flag[f], tag) < tag) args.fam_list[atomicAdd(args.fam_count, 1)] = f;
    }
#if LANE_INVERSE
    if (regular) {
        // The whole G x G grid of img1 at d = 0: its P x P centre is the template, its +-1 neighbours give the Jacobian
        // of src/algorithm.cpp:75-79.  The reference forms those coordinates in fp32, fl32(fl32(kx + x) +- 1); with
        // kx + x exact (single coordinate family, checked above) that is fl32(kx + x +- 1), the grid's coordinate.
        // Image rows iyn .. iyn+G and columns ixn .. ixn+G.
        const int wx0 = (ixn - 2) & ~15, wy0 = iyn;
        const int ox = ixn - wx0;
        const int k = ox >> 2, sh = (ox & 3) * 8;
        const uint8_t *img1 = lv.base[0] + (size_t)img * lv.slot;
        float rowA[G + 1], rowB[G + 1];
        float Sa[G], Sb[G], Sc[G];  // sample rows r-2, r-1, r
        auto pixel_row = [&](int i, float (&row)[G + 1]) {
            const int ry = min(max(wy0 + i, 0), lv.rows - 1);
            const uint4 *rp = reinterpret_cast<const uint4 *>(img1 + (ptrdiff_t)ry * lv.pitch + wx0);
            row_from_regs(__ldg(rp), __ldg(rp + 1), k, sh, row);
        };
        pixel_row(0, rowA);
#pragma unroll
        for (int r = 0; r < G; ++r) {
            pixel_row(r + 1, rowB);
#pragma unroll
            for (int g = 0; g < G; ++g)
                Sc[g] = bilerp(omx[g], xx[g], omy[r], yy[r], rowA[g], rowA[g + 1], rowB[g], rowB[g + 1]);
            if (r >= 2) {
                const int y = r - 2;
#pragma unroll
                for (int x = 0; x < P; ++x) {
                    mine_f[y * kTplPitch + x] = Sb[x + 1];                                          // :65  I1
                    mine_f[kI1Count + y * kTplPitch + x] = __fadd_rn(Sb[x + 2], -Sb[x]);           // :76-77
                    mine_f[2 * kI1Count + y * kTplPitch + x] = __fadd_rn(Sc[x + 1], -Sa[x + 1]);   // :78-79
                }
            }
#pragma unroll
            for (int g = 0; g <= G; ++g) rowA[g] = rowB[g];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                Sa[g] = Sb[g];
                Sb[g] = Sc[g];
            }
        }
        mine_f[kRecCount] = 1.f;  // regularity flag
    }
#else
    if (regular) {
        // the PxP centre of the GxG grid needs image rows iyn+1 .. iyn+P+1 and columns ixn+1 .. ixn+P+1
        const int wx0 = (ixn - 2) & ~15, wy0 = iyn + 1;
        const int ox = ixn - wx0;
        const int k = ox >> 2, sh = (ox & 3) * 8;
        const uint8_t *img1 = lv.base[0] + (size_t)img * lv.slot;
        if (LANE_TPL_EARLY != 2) {
#pragma unroll
            for (int i = 0; i < kTplRows; ++i) {  // all loads in flight; rows clamped, columns from the aprons
                const int ry = min(max(wy0 + i, 0), lv.rows - 1);
                const uint4 *rp = reinterpret_cast<const uint4 *>(img1 + (ptrdiff_t)ry * lv.pitch + wx0);
                q[i][0] = __ldg(rp);
                q[i][1] = __ldg(rp + 1);
            }
        }   // (else: requested at the top -- `regular` implies `early`, and ixn / iyn are the origin used there)
        float rowA[G + 1], rowB[G + 1];
        row_from_regs(q[0][0], q[0][1], k, sh, rowA);
#pragma unroll
        for (int y = 0; y < P; ++y) {
            row_from_regs(q[y + 1][0], q[y + 1][1], k, sh, rowB);
#pragma unroll
            for (int x = 0; x < P; ++x) {
                const float v = bilerp(omx[x + 1], xx[x + 1], omy[y + 1], yy[y + 1], rowA[x + 1], rowA[x + 2], rowB[x + 1],
                                       rowB[x + 2]);
                if (kDirect) mine_f[y * kTplPitch + x] = v;
                else buf[y * kTplPitch + x] = v;
            }
#pragma unroll
            for (int g = 0; g <= G; ++g) rowA[g] = rowB[g];
        }
        if (kDirect) mine_f[kI1Count] = 1.f;  // regularity flag
        else buf[kI1Count] = 1.f;
    }
#endif
    if (!kDirect) {
        float4 *mine = &stage[warp][lane * (kTplStride / 4)];
#pragma unroll
        for (int i = 0; i < kTplStride / 4; ++i) mine[i] = make_float4(buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]);
    }
    __syncwarp();
    // the warp's first item is feature args.f0 + (idx - lane) of this level; n_warp of its items are valid
    const int idx0 = idx - lane;
    const int n_warp = (level >= 0) ? min(32, args.n_total - idx0) : 0;
    if (n_warp > 0) {
        float4 *out = reinterpret_cast<float4 *>(args.templates + ((size_t)level * args.tpl_features + (size_t)(args.f0 + idx0)) * kTplStride);
#pragma unroll
        for (int k = 0; k < kTplStride / 4; ++k) {
            const int e = k * 32 + lane;
            if (e < n_warp * (kTplStride / 4)) out[e] = stage[warp][e];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Solver kernel.
// ------------------------------------------------------------------------------------------------
#ifndef LANE_MAXREG
#define LANE_MAXREG 0
#endif
// Two instances, launched concurrently on two streams.  FAM = false: the features whose levels all have a single
// coordinate family per axis (every integer keypoint; the template kernel sorts the others out) -- no family code at
// all in this instance.  FAM = true: the features of the template kernel's family list.  A level on which an axis has
// two coordinate families (sub-pixel source keypoints near a power of two, see axis_families) takes one trip per family
// combination for each pass, with the pixels of the other families AND-ed to zero; warps without such a thread in a
// trip run the unmasked row loop.  Measured on 512,000 sub-pixel keypoints (tools/ab.sh): one merged instance costs
// the integer-keypoint path 4 % (1.51 vs 1.45 ms) and tracks sub-pixel keypoints in 1.93 ms.
template <int T, int MIN_CTAS, bool FAM>
__global__ void
#if LANE_MAXREG > 0
__maxnreg__(LANE_MAXREG)
#else
__launch_bounds__(T, MIN_CTAS)
#endif
klt_lane_kernel(const __grid_constant__ PyramidView pyr, const __grid_constant__ SolverArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LaneSmem<T> &sm = *reinterpret_cast<LaneSmem<T> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned FULL = 0xffffffffu;
    int q_head = 0, q_tail = 0;  // warp-uniform ring positions
    constexpr bool second_phase = FAM;  // (this instance drains the family list)
    bool global_done = false;           // warp-uniform: no work is left

    if (tid < kStatCount) sm.stats[tid] = 0u;
    __syncthreads();

    const int L = pyr.levels;
    const double scale_top = 1.0 / (double)(1 << (L - 1));

    // ---- per-thread (per-feature) state ----
    int state = ST_FETCH, feat = -1, level = 0, iter = 0;
    float2 k1 = make_float2(0.f, 0.f), k2 = make_float2(0.f, 0.f);
    float kx = 0.f, ky = 0.f;
    double dx = 0, dy = 0, lastCost = 0;
    bool succ = true, flag = true;
    // Per-thread counters, flushed once when the thread runs out of work (atomics at every finished feature cost ~4 %
    // of the kernel: a divergent section with two or three active lanes).  Passes per level: four 16-bit fields per
    // word (levels 0-3, 4-7), flushed early before a field can overflow.
    unsigned long long it_lo = 0ull, it_hi = 0ull;
    unsigned n_nan = 0u, n_succ = 0u, n_out = 0u, n_feat_done = 0u;
    auto flush_counters = [&]() {
        for (int l = 0; l < L; ++l) {
            const unsigned v = (unsigned)(((l < 4 ? it_lo : it_hi) >> (16 * (l & 3))) & 0xffffull);
            if (v) atomicAdd(&sm.stats[kStatIters0 + l], v);
        }
        if (n_nan) atomicAdd(&sm.stats[kStatNan], n_nan);
        if (n_succ) atomicAdd(&sm.stats[kStatSuccess], n_succ);
        if (n_out) atomicAdd(&sm.stats[kStatOutOfImage], n_out);
        it_lo = it_hi = 0ull;
        n_nan = n_succ = n_out = n_feat_done = 0u;
    };
    int img = 0;
    int wx0 = 0, wy0 = 0;
    bool need_win = false, no_window = false;
    // coordinate families of this level: bits 0-6 x mask of family B, 8-14 y mask, 16-17 current sub-pass
    unsigned fam = 0u;
    double H00s = 0, H01s = 0, H11s = 0;  // inverse mode: the normal matrix of the level's first pass (src/algorithm.cpp:59,85-87)
#if !LANE_PARK_SMEM
    double pk0 = 0, pk1 = 0, pk2 = 0, pk3 = 0, pk4 = 0, pk5 = 0;  // partial sums between the sub-passes of a pass
#endif

    for (;;) {
        // ------------------------------------------------------------------ fetch new features
        {
            const unsigned m = __ballot_sync(FULL, state == ST_FETCH);
            if (m) {
                const int need = __popc(m);
                int avail = q_tail - q_head;
                if (avail < need && !global_done) {
                    // refill: top the ring up with consecutive feature ids (never more than its free slots),
                    // coalesced loads, compacted into the ring
                    const int want = min(32, kQueue - avail);
                    const int n_work = second_phase ? *args.fam_count : args.n_total;
                    int base = 0;
                    if (lane == 0) base = atomicAdd(args.work_counter + (second_phase ? 3 : 0), want);
                    base = __shfl_sync(FULL, base, 0);
                    const int local = base + lane;
                    bool keep = lane < want && local < n_work;
                    float2 a1 = make_float2(0.f, 0.f), a2 = a1;
                    int gid = args.f0 + local;
                    if (keep) {
                        if (second_phase) {
                            gid = args.fam_list[local];
                            keep = args.feat_flag[gid] == 4 * args.epoch + 1;  // (not raised to the deferred class)
                        } else {
                            keep = args.feat_flag[gid] < 4 * args.epoch;       // first phase: single-family features
                        }
                        a1 = args.kp1[gid];
                        a2 = args.kp2_init[gid];
                        if (slot_unused(args, gid, gid / args.n_per_pair)) {  // ragged batch: nothing to track here
                            write_unused_slot(args, gid);
                            keep = false;
                        }
                    }
                    const unsigned km = __ballot_sync(FULL, keep);
                    if (keep) {
                        const int pos = (q_tail + __popc(km & ((1u << lane) - 1u))) & (kQueue - 1);
                        sm.q_k1[warp][pos] = a1;
                        sm.q_k2[warp][pos] = a2;
                        sm.q_id[warp][pos] = gid;
                    }
                    q_tail += __popc(km);
                    if (base + want >= n_work) global_done = true;
                    __syncwarp();
                    avail = q_tail - q_head;
                }
                if (state == ST_FETCH) {
                    const int rank = __popc(m & ((1u << lane) - 1u));
                    if (rank < avail) {
                        const int pos = (q_head + rank) & (kQueue - 1);
                        feat = sm.q_id[warp][pos];
                        img = feat / args.n_per_pair;
                        k1 = sm.q_k1[warp][pos];
                        k2 = sm.q_k2[warp][pos];
                        k1.x = (float)(k1.x * scale_top);  // src/algorithm.cpp:160-169
                        k1.y = (float)(k1.y * scale_top);
                        k2.x = (float)(k2.x * scale_top);
                        k2.y = (float)(k2.y * scale_top);
                        level = L - 1;
                        flag = true;
                        state = ST_LEVEL;
                    } else if (global_done) {
                        state = ST_DONE;
                        flush_counters();
                    }
                }
                q_head += min(need, avail);
                __syncwarp();
            }
            if (__all_sync(FULL, state == ST_DONE)) break;
        }

        // ------------------------------------------------------------------ batched per-thread set-up
        {
            const bool blocked = (state == ST_LEVEL) || (state == ST_RUN && need_win);
            const int n_blocked = __popc(__ballot_sync(FULL, blocked));
            const int n_runnable = __popc(__ballot_sync(FULL, state == ST_RUN && !need_win));
            if (n_blocked > 0 && (n_blocked >= kSetupBatch || n_runnable == 0) && blocked) {
                const LevelView &lv = pyr.lv[level];
                const bool new_level = (state == ST_LEVEL);
                if (new_level) {
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
                    fam = 0u;
                    if (FAM) {
                        unsigned mBx, mBy;
                        double e0, e1;
                        axis_families(kx, mBx, e0);  // (more than one eps per axis was filtered by the template kernel)
                        axis_families(ky, mBy, e1);
                        fam = mBx | (mBy << kFamY);
                    }
                }
                const double Sx = (double)kx + dx, Sy = (double)ky + dy;
                no_window = true;  // estimate far outside the image / its apron: exact per-pixel passes
                if (fabs(Sx) < 1.0e6 && fabs(Sy) < 1.0e6) {
                    const int ixn = __double2int_rd(Sx + (double)(LO - 1)), iyn = __double2int_rd(Sy + (double)(LO - 1));
                    wx0 = (ixn - kWinSlackL) & ~(kWinAlign - 1);
                    wy0 = iyn - kWinSlackT;
                    if ((kWinSpare & 1) && Sy + (double)(LO - 1) - (double)iyn < 0.5) --wy0;  // odd spare row: on the nearer side
                    no_window = !window_in_apron(lv, wx0);
                }
                // Everything the level needs goes global -> shared memory by 16-byte cp.async copies (template record,
                // two per window row), in flight together; they land while the grid coordinates of the pass are
                // computed and are waited for just before the pass.
                const uint8_t *img2 = lv.base[1] + (size_t)img * lv.slot;
                if (new_level) {
                    const char *tp = reinterpret_cast<const char *>(args.templates + ((size_t)level * args.tpl_features + (size_t)feat) * kTplStride);
                    if (level > 0) {  // the next level's template will be needed a few trips from now
                        const char *nxt = tp - args.tpl_features * (kTplStride * sizeof(float));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128));
                    }
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.rec[0][tid]);
#pragma unroll
                    for (int j = 0; j < kRecCount / 4; ++j)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * T * 16), "l"(tp + 16 * j) : "memory");
                }
                if (!no_window) {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.win[0][tid][0]);
                    constexpr int kUnitBytes = 4 * kWinUnit, kUPR = LaneSmem<T>::kUnitsPerRow;
#pragma unroll
                    for (int i = 0; i < kWin2Rows; ++i) {
                        const int ry = min(max(wy0 + i, 0), lv.rows - 1);
                        const uint8_t *rp = img2 + (ptrdiff_t)ry * lv.pitch + wx0;
#pragma unroll
                        for (int j = 0; j < kUPR; ++j) {
                            const uint32_t d = dst + (i * kUPR + j) * T * kUnitBytes;
                            if (kWinUnit == 4)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(rp + 16 * j) : "memory");
                            else if (kWinUnit == 2)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(rp + 8 * j) : "memory");
                            else
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(rp + 4 * j) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                need_win = false;
                state = ST_RUN;
            }
        }

        // ------------------------------------------------------------------ grid coordinates
        bool run = (state == ST_RUN) && !need_win;
        bool fast = false;
        float xx[G], omx[G];
        int ixn = 0, iyn = 0;
        unsigned pmx = kPMask, pmy = kPMask;  // pixels (columns / rows) that belong to this sub-pass
        bool masked = false;                  // this trip covers only some of the pixels (pmx / pmy)
        Axis yax;
        if (run) {
            const LevelView &lv = pyr.lv[level];
            double kxd = (double)kx, kyd = (double)ky;
            if (FAM && (fam & kPMask2)) {  // a level with two coordinate families on some axis: this sub-pass works on one
                                   // (x family, y family) combination.  (Measured: the shift between the families changes a
                                   // rounded sample coordinate in 93 % of such passes -- no point in testing for it.)
                const unsigned mBx = fam & kPMask, mBy = (fam >> kFamY) & kPMask, sub = (fam >> kFamSub) & 3u;
                if (sub == 0u) atomicAdd(&sm.stats[kStatFamPasses], 1u);
                masked = true;
                const bool fx = mBx && (sub & 1u), fy = mBy && (mBx ? (sub >> 1) : (sub & 1u));
                if (fx) {
                    const int c = LO + __ffs(mBx) - 1;
                    kxd += (double)(kx + (float)c) - (kxd + (double)c);
                }
                if (fy) {
                    const int c = LO + __ffs(mBy) - 1;
                    kyd += (double)(ky + (float)c) - (kyd + (double)c);
                }
                pmx = fx ? mBx : (~mBx & kPMask);
                pmy = fy ? mBy : (~mBy & kPMask);
            }
            const bool okx = grid_axis<false>(kxd, dx, lv.cols, ixn, xx, omx);
            // rows: only the origin here; the row factors are formed inside the row loop of the pass (axis_factor),
            // their condition (b) is known after the loop
            yax = axis_begin(kyd, dy);
            iyn = yax.origin;
            if (!(okx && yax.in_range)) {
                // cannot prove the shared grid bit-identical for this pass (rare): exact per-pixel pass
                atomicAdd(&sm.stats[kStatSlowPath], 1u);
            } else if (!(ixn >= wx0 && ixn + (G + 1) <= wx0 + kWin2Words * 4 && iyn >= wy0 && iyn + (G + 1) <= wy0 + kWin2Rows)) {
                if (no_window) {
                    atomicAdd(&sm.stats[kStatSlowPath], 1u);
                } else {
                    need_win = true;  // footprint drifted out of the staged window: wait for the next set-up
                    run = false;
                }
            } else if (no_window) {
                atomicAdd(&sm.stats[kStatSlowPath], 1u);
            } else {
                fast = true;
            }
        }

        // ------------------------------------------------------------------ one Gauss-Newton pass
        asm volatile("cp.async.wait_all;" ::: "memory");  // this thread's set-up copies (each thread reads only its own granules)
        const bool any_masked = FAM && __any_sync(FULL, run && fast && masked);
        if (run) {
            const LevelView &lv = pyr.lv[level];
            const float *i1p = reinterpret_cast<const float *>(&sm.rec[0][tid]);
            constexpr int kI1Stride = T * 4;   // floats between consecutive granules of one thread
            double sb0 = 0, sb1 = 0, sc = 0, s00 = 0, s01 = 0, s11 = 0;
            bool solve_now = true;
            if (fast) {
            const int ox = ixn + kG0 - wx0;   // (inverse mode samples grid indices 1..P only)
            const int sh = (ox & 3) * 8;
            WinRef wr;
            {   // word kw + k of a row: copy unit (kw + k) / kWinUnit of the row, position (kw + k) % kWinUnit inside it
                const int kw = ox >> 2;
                constexpr int kUPR = LaneSmem<T>::kUnitsPerRow;
                wr.base = &sm.win[kUPR * (iyn + kG0 - wy0)][tid][0];
                wr.row_stride = kUPR * T * kWinUnit;
#pragma unroll
                for (int k = 0; k < kRowWords; ++k) wr.off[k] = ((kw + k) / kWinUnit) * (T * kWinUnit) + ((kw + k) % kWinUnit);
            }
            // Row factors of grid row r, formed when the row is sampled: n_r = (double)(LO - 1 + r) and
            // fg_r = (float)(origin + r) are carried through the loop (both exact).
            const float frows = (float)lv.rows, flastrow = (float)(lv.rows - 1);
            double n_r = (double)(LO - 1 + kG0);
            float fg_r = __fadd_rn(yax.forigin, (float)kG0);
            unsigned near_y = 0xFFFFFFFFu;
            auto row_weights = [&](float &om, float &fr) {
                axis_factor<true>(yax, n_r, fg_r, frows, flastrow, near_y, om, fr);
                n_r += 1.0;
                fg_r = __fadd_rn(fg_r, 1.f);
            };
            auto template_row = [&](int y, float (&v)[P]) {          // kTplPitch / 4 granules per row (y: row of any plane)
#pragma unroll
                for (int q = 0; q < kTplPitch / 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4 *>(i1p + (y * (kTplPitch / 4) + q) * (T * 4));
                    if (4 * q < P) v[4 * q] = t.x;
                    if (4 * q + 1 < P) v[4 * q + 1] = t.y;
                    if (4 * q + 2 < P) v[4 * q + 2] = t.z;
                    if (4 * q + 3 < P) v[4 * q + 3] = t.w;
                }
            };
            Row2 rowA, rowB;
            float2 Sa[kNP], Sb[kNP], Sc[kNP];  // sample rows r-2, r-1, r as pairs (2j, 2j+1); rotated by register moves
            float2 OMX[kNP], XX[kNP];
#pragma unroll
            for (int j = 0; j < kNP; ++j) {
                OMX[j] = make_float2(omx[kG0 + 2 * j], 2 * j + 1 < kNS ? omx[kG0 + 2 * j + 1] : 0.f);
                XX[j] = make_float2(xx[kG0 + 2 * j], 2 * j + 1 < kNS ? xx[kG0 + 2 * j + 1] : 0.f);
            }
#if LANE_INVERSE
            // Inverse mode: one img2 sample per pixel (its centre); the gradient differences come from the record --
            // the img1 gradients of the first pass, or, in later passes, the last pixel's copied over all pixels.
            load_row10_packed(wr, 0, sh, rowA);
#pragma unroll 1
            for (int y = 0; y < P; ++y) {
                load_row10_packed(wr, y + 1, sh, rowB);
                float om_s, fr_s, tpl[P], rgx[P], rgy[P];
                row_weights(om_s, fr_s);
                template_row(y, tpl);
                template_row(P + y, rgx);
                template_row(2 * P + y, rgy);
                sample_row_packed(OMX, XX, om_s, fr_s, rowA, rowB, args.one, Sc);
#pragma unroll
                for (int x = 0; x < P; ++x) {
                    const double e = (double)__fadd_rn(tpl[x], -pick(Sc, x));  // :65-66
                    const double gx = (double)rgx[x], gy = (double)rgy[x];     // :75-79 (first pass) / stale J (:57)
                    sb0 = fma(e, gx, sb0);
                    sb1 = fma(e, gy, sb1);
                    sc = fma(e, e, sc);
                    s00 = fma(gx, gx, s00);
                    s01 = fma(gx, gy, s01);
                    s11 = fma(gy, gy, s11);
                }
                rowA = rowB;
            }
#else

            // The row loop is deliberately NOT unrolled: the unrolled pass (1785 SASS instructions) did
            // not fit the instruction cache and the kernel was fetch-bound (profiles/README.md).
            float om_r, fr_r;
            load_row10_packed(wr, 0, sh, rowA);
            load_row10_packed(wr, 1, sh, rowB);
            row_weights(om_r, fr_r);
            sample_row_packed(OMX, XX, om_r, fr_r, rowA, rowB, args.one, Sa);
            rowA = rowB;
            load_row10_packed(wr, 2, sh, rowB);
            row_weights(om_r, fr_r);
            sample_row_packed(OMX, XX, om_r, fr_r, rowA, rowB, args.one, Sb);
            rowA = rowB;
            // The row loop, in two versions chosen per warp and trip: MASKED (some thread of the warp works on one family
            // combination of a two-family level: its other pixels are predicated off) and plain.
            auto rows = [&](auto masked_c) {
            constexpr bool MASKED = decltype(masked_c)::value;
            unsigned col_word[P];
#pragma unroll
            for (int x = 0; x < P; ++x) col_word[x] = (MASKED && !((pmx >> x) & 1u)) ? 0u : 0xFFFFFFFFu;
            // One patch row per step: sample row r from pixel rows r, r+1 (PB is loaded here), then the 7 pixels of
            // patch row y = r-2 with centre samples SB = grid row r-1, SA / SC the rows above / below.
            auto step = [&](int r, const Row2 &PA, Row2 &PB, const float2 (&SA)[kNP], const float2 (&SB)[kNP], float2 (&SC)[kNP]) {
                load_row10_packed(wr, r + 1, sh, PB);
                float om_s, fr_s, tpl[P];
                row_weights(om_s, fr_s);
                template_row(r - 2, tpl);
                sample_row_packed(OMX, XX, om_s, fr_s, PA, PB, args.one, SC);
                // MASKED: the three differences of a pixel outside this thread's sub-pass are AND-ed to +0 (one LOP3
                // each: difference & column word & row word), so that the pixel adds exact zeros to every sum
                const unsigned row_word = (!MASKED || ((pmy >> (r - 2)) & 1u)) ? 0xFFFFFFFFu : 0u;
#pragma unroll
                for (int x = 0; x < P; ++x) {
                    const int g = x + 1;
                    float ef = __fadd_rn(tpl[x], -pick(SB, g));                // :65-66
                    float gxf = __fadd_rn(pick(SB, g + 1), -pick(SB, g - 1));  // :70-71
                    float gyf = __fadd_rn(pick(SC, g), -pick(SA, g));          // :72-73
                    if (MASKED) {
                        ef = __uint_as_float(__float_as_uint(ef) & col_word[x] & row_word);
                        gxf = __uint_as_float(__float_as_uint(gxf) & col_word[x] & row_word);
                        gyf = __uint_as_float(__float_as_uint(gyf) & col_word[x] & row_word);
                    }
                    const double e = (LANE_WIDEN_ALU >= 1) ? widen_alu(ef) : (double)ef,
                                 gx = (LANE_WIDEN_ALU >= 2) ? widen_alu(gxf) : (double)gxf,
                                 gy = (LANE_WIDEN_ALU >= 3) ? widen_alu(gyf) : (double)gyf;
                    sb0 = fma(e, gx, sb0);
                    sb1 = fma(e, gy, sb1);
                    sc = fma(e, e, sc);
                    s00 = fma(gx, gx, s00);
                    s01 = fma(gx, gy, s01);
                    s11 = fma(gy, gy, s11);
                }
            };
            // Rolled on purpose: the fully unrolled pass was instruction-fetch bound (profiles/README.md).
            // Two steps per trip: the pixel rows swap roles by name (no moves), the three sample rows are renamed with
            // one rotation per trip (2 * kNP float2 moves per two steps instead of 4 * kNP per step).
            int r = 2;
            if ((G - 2) & 1) {  // odd number of steps: the first one is peeled
                step(2, rowA, rowB, Sa, Sb, Sc);
                rowA = rowB;
#pragma unroll
                for (int j = 0; j < kNP; ++j) {
                    Sa[j] = Sb[j];
                    Sb[j] = Sc[j];
                }
                r = 3;
            }
#pragma unroll 1
            for (; r < G; r += 2) {
                step(r, rowA, rowB, Sa, Sb, Sc);      // pixel row r+1 -> rowB, sample row r -> Sc
                step(r + 1, rowB, rowA, Sb, Sc, Sa);  // pixel row r+2 -> rowA, sample row r+1 -> Sa
#pragma unroll
                for (int j = 0; j < kNP; ++j) {       // (older, newer) = (Sc, Sa) -> (Sa, Sb)
                    const float2 t = Sa[j];
                    Sa[j] = Sc[j];
                    Sb[j] = t;
                }
            }
            };
            if (FAM && any_masked) {
                if (lane == __ffs(__activemask()) - 1) atomicAdd(&sm.stats[kStatMaskedTrips], 1u);
                rows(std::true_type{});
            } else {
                rows(std::false_type{});
            }
#endif

            if (!axis_ok(near_y, yax.tie)) {  // condition (b) failed on some grid row (rare): redo the pass exactly
                fast = false;
                atomicAdd(&sm.stats[kStatSlowPath], 1u);
            }
            }
            if (!fast) {
                // the reference formulation covers the whole patch whatever its families: restart the pass
                double sums[6];
#if LANE_INVERSE
                exact_pass_inverse(lv.base[1] + (size_t)img * lv.slot, lv, i1p, kI1Stride, kx, ky, dx, dy, sums);
#else
                exact_pass(lv.base[1] + (size_t)img * lv.slot, lv, i1p, kI1Stride, kx, ky, dx, dy, sums);
#endif
                sb0 = sums[0];
                sb1 = sums[1];
                sc = sums[2];
                s00 = sums[3];
                s01 = sums[4];
                s11 = sums[5];
                fam &= kPMask2;
            } else if (FAM && masked) {
                // this trip covered one (x family, y family) combination; the partial sums wait in registers until the
                // last combination has been added
                const unsigned sub = (fam >> kFamSub) & 3u;
                const unsigned nsub = ((fam & kPMask) ? 2u : 1u) * ((fam & (kPMask << kFamY)) ? 2u : 1u);
#if LANE_PARK_SMEM
                if (sub > 0u) {  // add what the earlier combinations left
                    sb0 += sm.parked[0][tid];
                    sb1 += sm.parked[1][tid];
                    sc += sm.parked[2][tid];
                    s00 += sm.parked[3][tid];
                    s01 += sm.parked[4][tid];
                    s11 += sm.parked[5][tid];
                }
                solve_now = sub + 1u >= nsub;
                if (!solve_now) {
                    sm.parked[0][tid] = sb0;
                    sm.parked[1][tid] = sb1;
                    sm.parked[2][tid] = sc;
                    sm.parked[3][tid] = s00;
                    sm.parked[4][tid] = s01;
                    sm.parked[5][tid] = s11;
                }
#else
                if (sub > 0u) {  // add what the earlier combinations left
                    sb0 += pk0;
                    sb1 += pk1;
                    sc += pk2;
                    s00 += pk3;
                    s01 += pk4;
                    s11 += pk5;
                }
                solve_now = sub + 1u >= nsub;
                if (!solve_now) {
                    pk0 = sb0;
                    pk1 = sb1;
                    pk2 = sc;
                    pk3 = s00;
                    pk4 = s01;
                    pk5 = s11;
                }
#endif
                fam = (fam & kPMask2) | (solve_now ? 0u : ((sub + 1u) << kFamSub));
            }
            if (solve_now) {

            // J = -0.5 * g: rescaling the sums by exact powers of two commutes with every rounding
            const double b0 = 0.5 * sb0, b1 = 0.5 * sb1, cost = sc;
            if (kInverse) {  // H is accumulated in the first pass of a level only and never reset (:59,85-87)
                if (iter == 0) {
                    H00s = s00;
                    H01s = s01;
                    H11s = s11;
                    // from now on every pixel uses the Jacobian of the first pass's LAST pixel (:57,74-80,83): copy its
                    // two gradient differences over the gradient planes of this thread's record
                    float *rec = const_cast<float *>(i1p);
                    const int last = (P - 1) * kTplPitch + (P - 1);
                    const float gxl = rec[((kI1Count + last) >> 2) * kI1Stride + ((kI1Count + last) & 3)];
                    const float gyl = rec[((2 * kI1Count + last) >> 2) * kI1Stride + ((2 * kI1Count + last) & 3)];
                    const float4 vx = make_float4(gxl, gxl, gxl, gxl), vy = make_float4(gyl, gyl, gyl, gyl);
#pragma unroll 1
                    for (int q = 0; q < kI1Count / 4; ++q) {
                        *reinterpret_cast<float4 *>(rec + (kI1Count / 4 + q) * kI1Stride) = vx;
                        *reinterpret_cast<float4 *>(rec + (2 * (kI1Count / 4) + q) * kI1Stride) = vy;
                    }
                } else {
                    s00 = H00s;
                    s01 = H01s;
                    s11 = H11s;
                }
            }
            const double H00 = 0.25 * s00, H01 = 0.25 * s01, H11 = 0.25 * s11;
            ++iter;
            bool level_done = false;
            double u0, u1;
            ldlt2_solve(H00, H01, H11, b0, b1, u0, u1);          // :92-93
            if (not_finite(u0) || not_finite(u1)) {               // :94-100
                ++n_nan;
                succ = false;
                level_done = true;
            } else if (iter > 1 && cost > lastCost) {             // :102-104 (iter here is 1-based)
                level_done = true;
            } else {
                dx = __dadd_rn(dx, u0);                           // :107-110
                dy = __dadd_rn(dy, u1);
                lastCost = cost;
                succ = true;
                if (__dadd_rn(__dmul_rn(u0, u0), __dmul_rn(u1, u1)) < args.eps_sq) level_done = true;  // :113
                if (iter >= args.max_iters) level_done = true;
            }

            if (level_done) {
                if (level < 4) it_lo += (unsigned long long)iter << (16 * level);
                else it_hi += (unsigned long long)iter << (16 * (level - 4));
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
                    n_succ += flag ? 1u : 0u;
                    n_out += point_in_image(k2.x, k2.y, lv) ? 0u : 1u;
                    if (++n_feat_done >= 4000u) flush_counters();  // (<= 15 passes per level and feature: no field overflows)
                    state = ST_FETCH;
                }
            }
            }  // solve_now
        }
    }

    __syncthreads();
    if (tid < kStatCount && sm.stats[tid]) atomicAdd(&args.stats[tid], (unsigned long long)sm.stats[tid]);
}

constexpr int kLaneThreads = LANE_T;
constexpr int kLaneMinCtas = LANE_CTAS;
#ifndef LANE_TPL_T
#define LANE_TPL_T 128
#endif
constexpr int kTplThreads = LANE_TPL_T;

}  // namespace

bool LANE_FN(lane_kernel_supports)(const SolverArgs &args) {
    return args.patch_lo == LO && args.patch_hi == HI && (args.inverse != 0) == kInverse && args.max_iters >= 1 &&
           args.max_iters <= 15;
}

size_t LANE_FN(lane_template_bytes)(int n_total, int levels) { return (size_t)n_total * levels * kTplStride * sizeof(float); }

size_t LANE_FN(lane_scratch_bytes)(int sm_count) { return (size_t)sm_count * kLaneMinCtas * kLaneThreads * 6 * sizeof(double); }

cudaError_t LANE_FN(launch_klt_template)(const PyramidView &pyr, const SolverArgs &args, cudaStream_t stream) {
    if (args.n_total <= 0) return cudaSuccess;
    auto kernel = klt_template_kernel<kTplThreads>;
    const long long items = (long long)((args.n_total + 31) & ~31) * pyr.levels;
    const int grid = (int)((items + kTplThreads - 1) / kTplThreads);
    kernel<<<grid, kTplThreads, 0, stream>>>(pyr, args);
    note_launch();
    return cudaGetLastError();
}

// `stream`: the common instance (single-family features, args.f0 ..).  `families_stream`: the family instance, fed by
// the template kernel's fam_list.  It only depends on the template kernel, is launched FIRST and with a smaller grid:
// its CTAs take their share of the SMs at once (or exit at once if the list is empty -- every integer-keypoint batch),
// the common instance's persistent CTAs fill the rest and take over the SMs the family CTAs leave.
cudaError_t LANE_FN(launch_klt_lane)(const PyramidView &pyr, const SolverArgs &args, int sm_count, cudaStream_t stream,
                                     cudaStream_t families_stream) {
    if (args.n_total <= 0) return cudaSuccess;
#ifndef LANE_SMEM_PAD
#define LANE_SMEM_PAD 0        // occupancy experiments: extra dynamic shared memory per CTA (fewer CTAs per SM)
#endif
#ifndef LANE_GRID_CTAS
#define LANE_GRID_CTAS LANE_CTAS
#endif
    const size_t smem = sizeof(LaneSmem<kLaneThreads>) + LANE_SMEM_PAD;
    const int needed = (args.n_total + kLaneThreads - 1) / kLaneThreads;
    if (!kInverse) {  // (the inverse instance has no family handling: the template kernel defers those features)
        auto kernel = klt_lane_kernel<kLaneThreads, kLaneMinCtas, true>;
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        const int grid = sm_count * (LANE_FAM_CTAS < LANE_GRID_CTAS ? LANE_FAM_CTAS : LANE_GRID_CTAS);
        kernel<<<grid < needed ? grid : needed, kLaneThreads, smem, families_stream>>>(pyr, args);
        note_launch();
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    auto kernel = klt_lane_kernel<kLaneThreads, kLaneMinCtas, false>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const int grid = sm_count * LANE_GRID_CTAS;
    kernel<<<grid < needed ? grid : needed, kLaneThreads, smem, stream>>>(pyr, args);
    note_launch();
    return cudaGetLastError();
}

}  // namespace legoklt
