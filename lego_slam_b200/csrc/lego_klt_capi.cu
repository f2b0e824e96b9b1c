// lego_klt_capi.cu -- the C ABI of include/lego_klt.h: contexts, device-resident batches, uploads,
// kernel launches, downloads.  No CPU fallback exists behind any compute entry point.
//
// Reference boundary replaced: legoslam::LKOpticalFlow4Layer / LKOpticalFlow1Layer
// (include/legoslam/algorithm.h:123-136), called at src/frontend_g2o.cpp:473 and :515.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include "klt_kernels.h"

using namespace legoklt;

std::atomic<long long> legoklt::g_kernel_launches{0};

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(LEGO_KLT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                      \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// NVTX range around a host-side phase (ranges around K1 / K2 / H2D): the kernels and copies enqueued
// inside it are attributed to it by the profilers.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

enum { EV_START = 0, EV_H2D, EV_PYR, EV_SOLVE, EV_D2H, EV_COUNT };
constexpr int kRing = 64;  // per-run kernel timing ring (lego_klt_batch_timings)
constexpr size_t kIoStatsBytes = 256;                       // device counters (kStatCount x 8 bytes, padded)
constexpr size_t kIoHeadBytes = kIoStatsBytes + 256;        // + work counters (4 x kMaxChunks ints)
static_assert(kStatCount * sizeof(unsigned long long) <= kIoStatsBytes && 4 * 16 * sizeof(int) <= 256, "io head layout");
constexpr int kAutoLaneMinFeatures = 3000;  // LEGO_KLT_KERNEL_AUTO: LANE above, WARP up to this many features per call
                                            // (measured per call, solver only: 2000 features 0.19 / 0.20 ms WARP / LANE, 5000: 0.24 / 0.21)
constexpr int kMaxChunks = 16;  // chunks of the overlapped end-to-end path (lego_klt_track_batched)

}  // namespace

struct lego_klt_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    lego_klt_batch *single = nullptr;  // cached B=1 batch behind lego_klt_track / build_pyramid
    lego_klt_batch *single_b = nullptr;  // second one: the stereo half of lego_klt_track_frame
    uint8_t *pinned = nullptr;         // staging for the single-pair path
    size_t pinned_bytes = 0;
    uint8_t *d_tri = nullptr;          // triangulation scratch (grow-only)
    size_t tri_bytes = 0;
    uint8_t *d_gftt = nullptr;         // feature detection: workspace | image | mask | exclusion list | outputs (grow-only)
    size_t gftt_bytes = 0;
    int gftt_cols = 0, gftt_rows = 0;  // shape of the last detection (lego_klt_debug_read_eig)
    // image-upload staging: a ring of pinned slots, each guarded by an event, so that an upload does not have to
    // drain the stream before it may overwrite the staging memory (lego_klt_image_upload*)
    struct Stage {
        uint8_t *p = nullptr;
        size_t bytes = 0;
        cudaEvent_t done = nullptr;
        bool busy = false;
    } stage[3];
    int stage_next = 0;
    uint8_t *pin_io = nullptr;         // keypoint / result staging of the single-pair paths
    size_t pin_io_bytes = 0;
    // Ownership: batches and images keep a pointer to their context.  lego_klt_destroy with handles still alive only
    // marks the context; it is torn down when the last batch / image handle is destroyed (see include/lego_klt.h).
    int live_handles = 0;
    bool destroy_requested = false;
};

struct lego_klt_batch {
    lego_klt_ctx *ctx = nullptr;
    int B = 0, cols = 0, rows = 0, n_cap = 0, n_active = 0, levels = 0;
    size_t step = 0;
    PyramidPlan plan;
    PyramidView view;
    WarpKernelMaps *maps = nullptr;
    uint8_t *d_images = nullptr;  // all levels, both sets
    uint8_t *d_tight = nullptr;   // H2D landing buffer: 2 sets of B tight images
    // keypoints, flags and counters live in ONE allocation: [kp1 | kp2_init] and [stats | work | kp2_out | success],
    // so that the single-pair paths move each group with one copy (lego_klt_track / lego_klt_track_images)
    uint8_t *d_io = nullptr;
    size_t io_out_off = 0;         // byte offset of the output group in d_io
    float2 *d_kp1 = nullptr, *d_kp2_init = nullptr, *d_kp2_out = nullptr;
    uint8_t *d_success = nullptr;
    unsigned long long *d_stats = nullptr;
    unsigned long long *h_stats = nullptr;  // pinned
    cudaEvent_t ev[EV_COUNT] = {};
    cudaEvent_t ring[kRing][3] = {};  // run r: [0] before pyramid, [1] after pyramid, [2] after solver
    long long runs = 0;            // runs of any kind (epoch of the LANE ownership flags)
    long long timed_runs = 0;      // runs that recorded a slot of `ring` (lego_klt_batch_timings)
    bool lane_ready = false;       // every LANE-path allocation below exists
    int *d_work = nullptr;         // per chunk: [0] lane work counter, [1] deferred count, [2] family count,
                                   // [3] lane<FAMILIES> work counter
    uint8_t *d_detect = nullptr;   // batched feature detection: workspace | corners | scores | counts (grow-only)
    size_t detect_bytes = 0;
    int detect_max = 0;            // max_corners of the last batched detection (0: none yet)
    size_t detect_ws = 0;          // where its outputs start inside d_detect
    std::vector<int> detect_counts;
    int *d_defer_list = nullptr;   // [B * n_cap]
    int *d_fam_list = nullptr;     // [B * n_cap]
    float *d_templates = nullptr;  // LANE kernel: I1 patches, allocated on first use
    size_t templates_bytes = 0;
    const uint8_t *d_slot_valid = nullptr;  // per-slot mask in force (lego_klt_track_frame), not owned
    int *d_pair_count = nullptr;   // ragged batches: valid features per pair (lego_klt_batch_set_feature_counts)
    std::vector<int> h_pair_count;
    bool ragged = false;
    int pipeline_chunks = 0;       // lego_klt_batch_set_pipeline_chunks (0 = default)
    unsigned long long n_valid = 0;  // sum of the counts
    int *d_feat_flag = nullptr;    // LANE path: per-feature 'handed to the warp kernel' flag
    double *d_scratch = nullptr;   // LANE kernel: per-thread partial sums of multi-family levels
    cudaStream_t side = nullptr;   // deferred features run here, concurrently with the lane kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t copy = nullptr;   // chunked end-to-end path: H2D of chunk c+1 overlaps compute of chunk c
    cudaEvent_t ev_chunk[kMaxChunks] = {};
    cudaEvent_t ev_compute_done = nullptr;
    cudaStream_t d2h = nullptr;    // chunked end-to-end path: results of chunk c leave while chunk c+1 computes
    cudaEvent_t ev_done[kMaxChunks] = {};
    cudaEvent_t ev_join_c[kMaxChunks] = {};  // side-stream work of chunk c (deferred features, FAMILIES instance) done
    bool uploaded = false, ran = false, pyramids_valid = false, last_chunked = false;
    bool in_flight = false;        // lego_klt_track_batched_begin without its _end yet
    bool counted = false;          // created through lego_klt_batch_create (counts in ctx->live_handles)
    bool last_timed = false;       // the last run recorded a ring slot
    lego_klt_params last_params;
};

struct lego_klt_image {
    lego_klt_ctx *ctx = nullptr;
    int cols = 0, rows = 0, levels = 0;
    size_t step = 0;
    PyramidPlan plan;
    PyramidView view;              // base[0] == base[1] == this image's levels
    WarpKernelMaps *maps = nullptr;  // TMA descriptors for the role "img2"
    uint8_t *d_levels = nullptr;
    uint8_t *d_tight = nullptr;
    uint8_t *d_full = nullptr;     // landing buffer of lego_klt_image_upload_fullres (grow-only)
    size_t full_bytes = 0;
    bool valid = false;
    bool counted = false;          // counts in ctx->live_handles
};

namespace {

void ctx_teardown(lego_klt_ctx *ctx);

void ctx_release_handle(lego_klt_ctx *ctx) {
    if (--ctx->live_handles == 0 && ctx->destroy_requested) ctx_teardown(ctx);
}

void ctx_teardown(lego_klt_ctx *ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->d_tri) cudaFree(ctx->d_tri);
    if (ctx->d_gftt) cudaFree(ctx->d_gftt);
    for (auto &sl : ctx->stage) {
        if (sl.p) cudaFreeHost(sl.p);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    if (ctx->pin_io) cudaFreeHost(ctx->pin_io);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int validate_params(const lego_klt_params *p, int levels_of_batch) {
    if (!p) return fail(LEGO_KLT_ERR_BAD_ARG, "params is null");
    if (p->levels != levels_of_batch)
        return fail(LEGO_KLT_ERR_BAD_ARG, "params->levels (%d) != batch levels (%d)", p->levels, levels_of_batch);
    if (p->patch_lo > p->patch_hi || p->patch_hi - p->patch_lo + 1 > kMaxPatch)
        return fail(LEGO_KLT_ERR_UNSUPPORTED, "patch %d..%d unsupported (max %d wide)", p->patch_lo, p->patch_hi,
                    kMaxPatch);
    if (p->max_iters < 0) return fail(LEGO_KLT_ERR_BAD_ARG, "max_iters < 0");
    if (p->kernel < LEGO_KLT_KERNEL_AUTO || p->kernel > LEGO_KLT_KERNEL_PATCH)
        return fail(LEGO_KLT_ERR_BAD_ARG, "unknown kernel id %d", p->kernel);
    return LEGO_KLT_OK;
}

int batch_alloc(lego_klt_ctx *ctx, int B, int cols, int rows, size_t step, int n, int levels,
                lego_klt_batch **out) {
    if (!ctx || !out) return fail(LEGO_KLT_ERR_BAD_ARG, "null ctx/out");
    if (B <= 0 || cols <= 0 || rows <= 0 || step < (size_t)cols || n < 0 || levels < 1 || levels > kMaxLevels)
        return fail(LEGO_KLT_ERR_BAD_ARG, "bad batch shape B=%d cols=%d rows=%d step=%zu n=%d levels=%d", B, cols,
                    rows, step, n, levels);
    if (B > 32000) return fail(LEGO_KLT_ERR_UNSUPPORTED, "batch too large (%d > 32000)", B);
    int lc[kMaxLevels], lr[kMaxLevels];
    if (!pyramid_level_sizes(cols, rows, levels, lc, lr))
        return fail(LEGO_KLT_ERR_UNSUPPORTED, "pyramid level would be empty for %dx%d, %d levels", cols, rows,
                    levels);
    CU_TRY(cudaSetDevice(ctx->device));
    lego_klt_batch *b = new (std::nothrow) lego_klt_batch();
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "out of host memory");
    b->ctx = ctx;
    b->B = B;
    b->cols = cols;
    b->rows = rows;
    b->step = step;
    b->n_cap = n;
    b->n_active = n;
    b->levels = levels;

    // ---- device image blob: per level, per set: B images of rows*pitch bytes ----
    int pitch[kMaxLevels];
    size_t total = 0, off[kMaxLevels][2];
    memset(&b->view, 0, sizeof(b->view));
    b->view.levels = levels;
    b->view.n_images = B;
    for (int l = 0; l < levels; ++l) {
        LevelView &lv = b->view.lv[l];
        lv.cols = lc[l];
        lv.rows = lr[l];
        lv.step = (l == 0) ? (int)step : lc[l];
        pitch[l] = kApronL + (int)align_up((size_t)lv.step, 16) + kApronR;
        lv.pitch = pitch[l];
        lv.slot = (unsigned long long)lr[l] * pitch[l];
        for (int s = 0; s < 2; ++s) {
            off[l][s] = total;  // base = off + kApronL; the pyramid kernel's band copies run 32 bytes past the set
            total += align_up((size_t)B * lv.slot + 2 * kApronL + 64, 256);
        }
    }
    total += 4096;
    auto cleanup = [&](int code) {
        lego_klt_batch_destroy(b);
        return code;
    };
    cudaError_t e = cudaMalloc(&b->d_images, total);
    if (e != cudaSuccess) return cleanup(fail(LEGO_KLT_ERR_CUDA, "cudaMalloc(%zu) images: %s", total, cudaGetErrorString(e)));
    e = cudaMemsetAsync(b->d_images, 0, total, ctx->stream);
    if (e != cudaSuccess) return cleanup(fail(LEGO_KLT_ERR_CUDA, "cudaMemset images: %s", cudaGetErrorString(e)));
    for (int l = 0; l < levels; ++l)
        for (int s = 0; s < 2; ++s) b->view.lv[l].base[s] = b->d_images + off[l][s] + kApronL;
    e = cudaMalloc(&b->d_tight, 2 * align_up((size_t)B * rows * step + 256, 256));
    if (e != cudaSuccess) return cleanup(fail(LEGO_KLT_ERR_CUDA, "cudaMalloc landing buffer: %s", cudaGetErrorString(e)));

    const size_t nt = (size_t)B * (size_t)(n > 0 ? n : 1);
    b->io_out_off = align_up(2 * nt * sizeof(float2), 256);
    const size_t io_bytes = b->io_out_off + kIoHeadBytes + nt * sizeof(float2) + align_up(nt, 256);
    if ((e = cudaMalloc(&b->d_io, io_bytes)) != cudaSuccess)
        return cleanup(fail(LEGO_KLT_ERR_CUDA, "allocating keypoint buffers: %s", cudaGetErrorString(e)));
    b->d_kp1 = reinterpret_cast<float2 *>(b->d_io);
    b->d_kp2_init = b->d_kp1 + nt;
    b->d_stats = reinterpret_cast<unsigned long long *>(b->d_io + b->io_out_off);
    b->d_work = reinterpret_cast<int *>(b->d_io + b->io_out_off + kIoStatsBytes);
    b->d_kp2_out = reinterpret_cast<float2 *>(b->d_io + b->io_out_off + kIoHeadBytes);
    b->d_success = reinterpret_cast<uint8_t *>(b->d_kp2_out + nt);
    if ((e = cudaMalloc(&b->d_defer_list, nt * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&b->d_fam_list, nt * sizeof(int))) != cudaSuccess ||
        (e = cudaMallocHost(&b->h_stats, kStatCount * sizeof(unsigned long long))) != cudaSuccess)
        return cleanup(fail(LEGO_KLT_ERR_CUDA, "allocating keypoint buffers: %s", cudaGetErrorString(e)));
    for (int i = 0; i < EV_COUNT; ++i)
        if ((e = cudaEventCreate(&b->ev[i])) != cudaSuccess)
            return cleanup(fail(LEGO_KLT_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e)));
    for (int r = 0; r < kRing; ++r)
        for (int i = 0; i < 3; ++i)
            if ((e = cudaEventCreate(&b->ring[r][i])) != cudaSuccess)
                return cleanup(fail(LEGO_KLT_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e)));
    if ((e = pyramid_plan_create(cols, rows, levels, pitch, &b->plan)) != cudaSuccess)
        return cleanup(fail(LEGO_KLT_ERR_CUDA, "pyramid plan: %s", cudaGetErrorString(e)));
    if ((e = warp_maps_create(b->view, &b->maps)) != cudaSuccess)
        return cleanup(fail(LEGO_KLT_ERR_CUDA, "TMA descriptor creation failed: %s", cudaGetErrorString(e)));
    *out = b;
    return LEGO_KLT_OK;
}

// H2D of level 0 of images [img0, img0+nimg) of one image set: a plain contiguous copy into the landing
// buffer on `copy_stream`, then (after `ready`, if given) the re-pitch kernel on `kernel_stream`.
cudaError_t upload_set(lego_klt_batch *b, int set, const uint8_t *src, int img0, int nimg, cudaStream_t copy_stream) {
    const size_t img_bytes = (size_t)b->rows * b->step;
    const size_t set_bytes = align_up((size_t)b->B * img_bytes + 256, 256);
    uint8_t *landing = b->d_tight + (size_t)set * set_bytes + (size_t)img0 * img_bytes;
    return cudaMemcpyAsync(landing, src + (size_t)img0 * img_bytes, (size_t)nimg * img_bytes, cudaMemcpyHostToDevice,
                           copy_stream);
}

cudaError_t ingest_set(lego_klt_batch *b, int set, int img0, int nimg, cudaStream_t stream) {
    const size_t img_bytes = (size_t)b->rows * b->step;
    const size_t set_bytes = align_up((size_t)b->B * img_bytes + 256, 256);
    // the funnel-shift loads need a 4-byte aligned base: pass the set base and let the kernel index rows from img0
    return launch_ingest(b->d_tight + (size_t)set * set_bytes, b->view.lv[0], set, img0, nimg, stream);
}

void free_lane_buffers(lego_klt_batch *b) {
    cudaFree(b->d_feat_flag);
    cudaFree(b->d_scratch);
    if (b->side) cudaStreamDestroy(b->side);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    b->d_feat_flag = nullptr;
    b->d_scratch = nullptr;
    b->side = nullptr;
    b->ev_fork = b->ev_join = nullptr;
    b->lane_ready = false;
    cudaGetLastError();
}

int ensure_lane_buffers(lego_klt_batch *b, size_t template_bytes) {
    // templates: sized for the patch in hand, re-allocated if a larger patch follows (both streams that read them
    // are drained first)
    if (template_bytes > b->templates_bytes) {
        CU_TRY(cudaStreamSynchronize(b->ctx->stream));
        if (b->side) CU_TRY(cudaStreamSynchronize(b->side));
        if (b->d_templates) cudaFree(b->d_templates);
        b->d_templates = nullptr;
        b->templates_bytes = 0;
        CU_TRY(cudaMalloc(&b->d_templates, template_bytes));
        b->templates_bytes = template_bytes;
    }
    if (b->lane_ready) return LEGO_KLT_OK;
    // all or nothing: a failure half way leaves no partly initialised state behind for the next call
    const size_t cap = (size_t)b->B * (size_t)(b->n_cap > 0 ? b->n_cap : 1);
    const size_t scratch = std::max(std::max(lane_scratch_bytes(b->ctx->sm_count), lane_scratch_bytes_p8(b->ctx->sm_count)),
                                    lane_scratch_bytes_p11(b->ctx->sm_count));
    cudaError_t e;
    if ((e = cudaMalloc(&b->d_feat_flag, cap * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&b->d_scratch, scratch)) != cudaSuccess ||
        (e = cudaMemset(b->d_feat_flag, 0, cap * sizeof(int))) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&b->side, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming)) != cudaSuccess) {
        free_lane_buffers(b);
        return fail(LEGO_KLT_ERR_CUDA, "allocating the LANE-path buffers: %s", cudaGetErrorString(e));
    }
    b->lane_ready = true;
    return LEGO_KLT_OK;
}

// Pyramids + aprons + solver for images [img0, img0+nimg) (features f0 = img0*n .. ), on the context stream.
// `chunk` selects the set of device work counters; `ring` (3 events) is recorded around the kernels if given.
int run_range(lego_klt_batch *b, const lego_klt_params *params, int img0, int nimg, int chunk, cudaEvent_t *ring,
              const PyramidView *view_override = nullptr, const WarpKernelMaps *maps_override = nullptr,
              cudaEvent_t deferred_join = nullptr) {
    lego_klt_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    const PyramidView &view = view_override ? *view_override : b->view;   // cached pyramids of image handles
    const WarpKernelMaps *maps = maps_override ? maps_override : b->maps;
    int *work = b->d_work + 4 * chunk;
    if (ring) CU_TRY(cudaEventRecord(ring[0], st));
    if (!view_override) {
        NvtxRange r("lego_klt K1 pyramid");
        CU_TRY(launch_pyramid(b->plan, b->view, img0, nimg, st));  // (row aprons included)
    }
    if (ring) CU_TRY(cudaEventRecord(ring[1], st));
    NvtxRange solver_range("lego_klt K2 solver");
    SolverArgs a;
    a.kp1 = b->d_kp1;
    a.kp2_init = b->d_kp2_init;
    a.kp2_out = b->d_kp2_out;
    a.success = b->d_success;
    a.stats = b->d_stats;
    a.n_per_pair = b->n_active > 0 ? b->n_active : 1;
    a.pair_count = b->ragged ? b->d_pair_count : nullptr;
    a.slot_valid = b->d_slot_valid;
    a.n_total = nimg * b->n_active;
    a.f0 = img0 * b->n_active;
    a.patch_lo = params->patch_lo;
    a.patch_hi = params->patch_hi;
    a.max_iters = params->max_iters;
    a.inverse = params->inverse;
    a.has_initial = params->has_initial;
    a.eps = params->eps;
    {   // threshold on the squared norm equivalent to `update.norm() < eps` (src/algorithm.cpp:113), see SolverArgs
        double c = params->eps * params->eps;
        if (params->eps > 0 && std::isfinite(c) && c > 0) {
            while (std::sqrt(std::nextafter(c, 0.0)) >= params->eps) c = std::nextafter(c, 0.0);
            while (std::sqrt(c) < params->eps) c = std::nextafter(c, INFINITY);
        } else {
            c = (params->eps > 0) ? INFINITY : 0.0;  // eps <= 0: never converged by the norm test
        }
        a.eps_sq = c;
    }
    a.one = 1.0f;
    {
        const char *dbg = getenv("LEGO_KLT_DEBUG");
        a.debug_flags = dbg ? atoi(dbg) : 0;
    }
    a.list = nullptr;
    a.list_count = nullptr;
    a.work_counter = work;
    a.defer_count = work + 1;
    a.defer_list = b->d_defer_list + a.f0;
    a.fam_list = b->d_fam_list + a.f0;
    a.fam_count = work + 2;
    a.templates = nullptr;
    a.tpl_features = 0;
    a.scratch = nullptr;
    a.feat_flag = nullptr;
    a.epoch = 0;
    int kernel = params->kernel;
    // AUTO: the thread-per-feature LANE kernel needs tens of thousands of features to fill the machine (47k resident
    // threads); below ~3k features of one call the warp-per-feature kernel has the lower latency (measured, 1241x376,
    // tools/seq_latency.py: n = 150: 0.13 vs 0.21 ms per call, 2000: 0.19 vs 0.20, 5000: 0.24 vs 0.21, 20000: 0.56 vs
    // 0.26).  Same fidelity contract.
    // the LANE solver is compiled for three patches (klt_solver_lane*.cu): 0 = none, else the patch width
    // (-7: the reference's inverse mode, 7x7)
    const int lane_patch = lane_kernel_supports(a)       ? 7
                           : lane_kernel_supports_p8(a)  ? 8
                           : lane_kernel_supports_p11(a) ? 11
                           : lane_kernel_supports_inv(a) ? -7
                                                         : 0;
    if (kernel == LEGO_KLT_KERNEL_AUTO)
        kernel = (lane_patch && a.n_total > kAutoLaneMinFeatures) ? LEGO_KLT_KERNEL_LANE
                 : (patch_kernel_supports(a) && a.n_total <= kAutoLaneMinFeatures) ? LEGO_KLT_KERNEL_PATCH
                                                                                   : LEGO_KLT_KERNEL_WARP;
    if (kernel == LEGO_KLT_KERNEL_LANE && !lane_patch)
        return fail(LEGO_KLT_ERR_UNSUPPORTED,
                    "LANE kernel: forward mode with the 7x7 (-3..3), 8x8 (-4..3) or 11x11 (-5..5) patch, or inverse mode "
                    "with the 7x7 patch");
    // the device work counters (feature queue, deferred / family lists) of the warp and lane kernels
    if (kernel == LEGO_KLT_KERNEL_WARP || kernel == LEGO_KLT_KERNEL_LANE) CU_TRY(cudaMemsetAsync(work, 0, 4 * sizeof(int), st));
    if (kernel == LEGO_KLT_KERNEL_EXACT) {
        CU_TRY(launch_klt_exact(view, a, st));
    } else if (kernel == LEGO_KLT_KERNEL_PATCH) {
        if (!patch_kernel_supports(a)) return fail(LEGO_KLT_ERR_UNSUPPORTED, "PATCH kernel: patches of up to 16 x 16");
        CU_TRY(launch_klt_patch(view, a, st));
    } else if (kernel == LEGO_KLT_KERNEL_WARP) {
        CU_TRY(launch_klt_warp(view, maps, a, ctx->sm_count, st));
    } else if (a.n_total > 0) {
        const size_t cap = (size_t)b->B * (size_t)(b->n_cap > 0 ? b->n_cap : 1);
        int rc = ensure_lane_buffers(b, lane_patch == 7    ? lane_template_bytes((int)cap, b->levels)
                                        : lane_patch == 8  ? lane_template_bytes_p8((int)cap, b->levels)
                                        : lane_patch == 11 ? lane_template_bytes_p11((int)cap, b->levels)
                                                           : lane_template_bytes_inv((int)cap, b->levels));
        if (rc) return rc;
        a.templates = b->d_templates;
        a.tpl_features = (unsigned long long)b->B * (unsigned long long)(b->n_cap > 0 ? b->n_cap : 1);
        a.feat_flag = b->d_feat_flag;
        a.scratch = b->d_scratch;
        a.epoch = (int)((b->runs % 0x0fffffff) + 1);
        if (a.epoch == 1 && b->runs > 0)  // the run counter wrapped: flags of earlier runs would look newer than this one
            CU_TRY(cudaMemsetAsync(b->d_feat_flag, 0, cap * sizeof(int), st));
        CU_TRY(lane_patch == 7    ? launch_klt_template(view, a, st)
               : lane_patch == 8  ? launch_klt_template_p8(view, a, st)
               : lane_patch == 11 ? launch_klt_template_p11(view, a, st)
                                  : launch_klt_template_inv(view, a, st));
        // features with an irregular template (kx+c inexact in fp32, ...) are solved by the exact warp
        // kernel on a second stream while the lane kernel solves the rest
        CU_TRY(cudaEventRecord(b->ev_fork, st));
        CU_TRY(cudaStreamWaitEvent(b->side, b->ev_fork, 0));
        SolverArgs aw = a;
        aw.list = a.defer_list;
        aw.list_count = a.defer_count;
        CU_TRY(launch_klt_warp(view, maps, aw, ctx->sm_count, b->side));
        // the family instance of the lane kernel (features with two coordinate families on some level) goes first, on
        // the side stream; the common instance on the context stream
        CU_TRY(lane_patch == 7    ? launch_klt_lane(view, a, ctx->sm_count, st, b->side)
               : lane_patch == 8  ? launch_klt_lane_p8(view, a, ctx->sm_count, st, b->side)
               : lane_patch == 11 ? launch_klt_lane_p11(view, a, ctx->sm_count, st, b->side)
                                  : launch_klt_lane_inv(view, a, ctx->sm_count, st, b->side));
        if (deferred_join) {
            CU_TRY(cudaEventRecord(deferred_join, b->side));   // the caller joins before it needs the results
        } else {
            CU_TRY(cudaEventRecord(b->ev_join, b->side));
            CU_TRY(cudaStreamWaitEvent(st, b->ev_join, 0));
        }
    }
    if (ring) CU_TRY(cudaEventRecord(ring[2], st));
    return LEGO_KLT_OK;
}

int batch_run(lego_klt_batch *b, const lego_klt_params *params) {
    int rc = validate_params(params, b->levels);
    if (rc) return rc;
    if (!b->uploaded) return fail(LEGO_KLT_ERR_STATE, "lego_klt_batch_run before upload");
    lego_klt_ctx *ctx = b->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaMemsetAsync(b->d_stats, 0, kStatCount * sizeof(unsigned long long), st));
    cudaEvent_t *ring = b->ring[b->timed_runs % kRing];
    CU_TRY(cudaEventRecord(b->ev[EV_H2D], st));
    rc = run_range(b, params, 0, b->B, 0, ring);
    if (rc) return rc;
    b->pyramids_valid = true;
    CU_TRY(cudaEventRecord(b->ev[EV_PYR], ctx->stream));  // kept for the stats struct; see fill_stats
    CU_TRY(cudaEventRecord(b->ev[EV_SOLVE], st));
    ++b->runs;
    ++b->timed_runs;
    b->last_timed = true;
    b->ran = true;
    b->last_params = *params;
    return LEGO_KLT_OK;
}

// Kernel times of run slot r: pyramid launches (incl. aprons) and everything else (templates + solver).
cudaError_t run_times(lego_klt_batch *b, int r, float *ms_pyr, float *ms_solver) {
    float pyr = 0.f, total = 0.f;
    cudaError_t e = cudaEventElapsedTime(&pyr, b->ring[r][0], b->ring[r][1]);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventElapsedTime(&total, b->ring[r][0], b->ring[r][2])) != cudaSuccess) return e;
    *ms_pyr = pyr;
    *ms_solver = total - pyr;
    return cudaSuccess;
}

void fill_stats(lego_klt_batch *b, lego_klt_stats *s) {
    memset(s, 0, sizeof(*s));
    s->n_features = b->ragged ? (uint64_t)b->n_valid : (uint64_t)b->B * (uint64_t)b->n_active;
    s->n_success = b->h_stats[kStatSuccess];
    s->n_nan = b->h_stats[kStatNan];
    s->n_out_of_image = b->h_stats[kStatOutOfImage];
    s->n_slow_path = b->h_stats[kStatSlowPath];
    s->n_deferred = b->h_stats[kStatDeferred];
    for (int i = 0; i < 4; ++i) s->defer_reason[i] = b->h_stats[kStatDeferInexact + i];
    if (b->h_stats[kStatTmaTimeout]) s->n_nan += 1000000ull * b->h_stats[kStatTmaTimeout];  // debug aid
    for (int l = 0; l < kMaxLevels; ++l) s->gn_iters[l] = b->h_stats[kStatIters0 + l];
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, b->ev[EV_START], b->ev[EV_H2D]) == cudaSuccess) s->ms_h2d = ms;
    if (b->timed_runs > 0 && !b->last_chunked && b->last_timed) {
        float mp = 0.f, msol = 0.f;
        if (run_times(b, (int)((b->timed_runs - 1) % kRing), &mp, &msol) == cudaSuccess) {
            s->ms_pyramid = mp;
            s->ms_solver = msol;
        }
    } else if (cudaEventElapsedTime(&ms, b->ev[EV_H2D], b->ev[EV_SOLVE]) == cudaSuccess) {
        s->ms_solver = ms;  // chunked run: copies and kernels overlap, only the total is meaningful
    }
    if (cudaEventElapsedTime(&ms, b->ev[EV_SOLVE], b->ev[EV_D2H]) == cudaSuccess) s->ms_d2h = ms;
    cudaGetLastError();  // events not recorded yet (run without upload) are not an error of this call
}

int ensure_pinned(lego_klt_ctx *ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return LEGO_KLT_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    CU_TRY(cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return LEGO_KLT_OK;
}

// Next slot of the upload staging ring, at least `bytes` large and no longer read by an earlier copy.
int acquire_stage(lego_klt_ctx *ctx, size_t bytes, lego_klt_ctx::Stage **out) {
    lego_klt_ctx::Stage &sl = ctx->stage[ctx->stage_next];
    ctx->stage_next = (ctx->stage_next + 1) % 3;
    if (!sl.done) CU_TRY(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    if (sl.busy) {
        CU_TRY(cudaEventSynchronize(sl.done));
        sl.busy = false;
    }
    if (sl.bytes < bytes) {
        if (sl.p) cudaFreeHost(sl.p);
        sl.p = nullptr;
        sl.bytes = 0;
        CU_TRY(cudaMallocHost(&sl.p, bytes));
        sl.bytes = bytes;
    }
    *out = &sl;
    return LEGO_KLT_OK;
}

int release_stage(lego_klt_ctx *ctx, lego_klt_ctx::Stage *sl) {  // call after the copy that reads it was enqueued
    CU_TRY(cudaEventRecord(sl->done, ctx->stream));
    sl->busy = true;
    return LEGO_KLT_OK;
}

// Single-pair paths: both keypoint arrays go up in ONE copy and counters + positions + flags come back in ONE
// copy, through pinned staging (caller memory is usually pageable: a cudaMemcpyAsync from / to it would make the
// driver stage and, for device-to-host, block the stream once per copy).  The cached batch's pointers are
// re-based for the call's n (groups packed, not at capacity stride).
int single_upload_keypoints(lego_klt_ctx *ctx, lego_klt_batch *b, const float *kp1_xy, const float *kp2_xy, int n) {
    const size_t kp_bytes = (size_t)n * sizeof(float2);
    const size_t need = 2 * kp_bytes + kIoHeadBytes + kp_bytes + (size_t)n + 64;
    if (ctx->pin_io_bytes < need) {
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->pin_io) cudaFreeHost(ctx->pin_io);
        ctx->pin_io = nullptr;
        ctx->pin_io_bytes = 0;
        CU_TRY(cudaMallocHost(&ctx->pin_io, need + need / 2));
        ctx->pin_io_bytes = need + need / 2;
    }
    b->n_active = n;
    b->d_kp2_init = b->d_kp1 + n;
    b->d_success = reinterpret_cast<uint8_t *>(b->d_kp2_out + n);
    if (n) {
        memcpy(ctx->pin_io, kp1_xy, kp_bytes);
        memcpy(ctx->pin_io + kp_bytes, kp2_xy, kp_bytes);
        CU_TRY(cudaMemcpyAsync(b->d_kp1, ctx->pin_io, 2 * kp_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    return LEGO_KLT_OK;
}

int single_download(lego_klt_ctx *ctx, lego_klt_batch *b, float *kp2_xy, uint8_t *success, int n, lego_klt_stats *stats) {
    const size_t kp_bytes = (size_t)n * sizeof(float2);
    uint8_t *h = ctx->pin_io + 2 * kp_bytes;  // (after the input group)
    h = reinterpret_cast<uint8_t *>(align_up((size_t)(uintptr_t)h, 16));
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaMemcpyAsync(h, b->d_io + b->io_out_off, kIoHeadBytes + kp_bytes + (size_t)n, cudaMemcpyDeviceToHost, st));
    if (stats) CU_TRY(cudaEventRecord(b->ev[EV_D2H], st));
    CU_TRY(cudaStreamSynchronize(st));
    memcpy(b->h_stats, h, kStatCount * sizeof(unsigned long long));
    if (n) {
        memcpy(kp2_xy, h + kIoHeadBytes, kp_bytes);
        memcpy(success, h + kIoHeadBytes + kp_bytes, (size_t)n);
    }
    if (stats) fill_stats(b, stats);
    return LEGO_KLT_OK;
}

// (Re)creates the cached single-pair batch when the shape changes or more features are needed.
int ensure_single(lego_klt_ctx *ctx, int cols, int rows, size_t step, int n, int levels, lego_klt_batch **slot = nullptr) {
    lego_klt_batch *&s = slot ? *slot : ctx->single;
    if (s && s->cols == cols && s->rows == rows && s->step == step && s->levels == levels && s->n_cap >= n)
        return LEGO_KLT_OK;
    if (s) lego_klt_batch_destroy(s);
    s = nullptr;
    int cap = n < 256 ? 256 : (int)align_up((size_t)n, 256);
    return batch_alloc(ctx, 1, cols, rows, step, cap, levels, &s);
}

}  // namespace

extern "C" {

int lego_klt_abi_version(void) { return LEGO_KLT_ABI_VERSION; }

long long lego_klt_kernel_launches(void) { return legoklt::g_kernel_launches.load(std::memory_order_relaxed); }

const char *lego_klt_last_error(void) { return g_last_error.c_str(); }

void lego_klt_default_params(lego_klt_params *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->levels = 4;        // src/algorithm.cpp:135
    p->patch_lo = -3;     // :40,63-64
    p->patch_hi = 3;
    p->max_iters = 10;    // :42
    p->inverse = 0;       // src/frontend_g2o.cpp:473,515
    p->has_initial = 1;
    p->kernel = LEGO_KLT_KERNEL_AUTO;
    p->eps = 1e-2;        // :113
}

int lego_klt_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(LEGO_KLT_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    }
    return n;
}

int lego_klt_create(int device, lego_klt_ctx **out) {
    if (!out) return fail(LEGO_KLT_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    int n = lego_klt_device_count();
    if (n < 0) return n;
    if (device < 0 || device >= n) return fail(LEGO_KLT_ERR_BAD_ARG, "device %d out of range [0,%d)", device, n);
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(LEGO_KLT_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device,
                    prop.major, prop.minor);
    lego_klt_ctx *ctx = new (std::nothrow) lego_klt_ctx();
    if (!ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(LEGO_KLT_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return LEGO_KLT_OK;
}

void lego_klt_destroy(lego_klt_ctx *ctx) {
    if (!ctx) return;
    if (ctx->single) {
        lego_klt_batch_destroy(ctx->single);
        ctx->single = nullptr;
    }
    if (ctx->single_b) {
        lego_klt_batch_destroy(ctx->single_b);
        ctx->single_b = nullptr;
    }
    if (ctx->live_handles > 0) {  // batches / images of the caller still point here: the last one tears down
        ctx->destroy_requested = true;
        return;
    }
    ctx_teardown(ctx);
}


int lego_klt_set_stream(lego_klt_ctx *ctx, void *cuda_stream) {
    if (!ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "ctx is null");
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return LEGO_KLT_OK;
}

int lego_klt_sync(lego_klt_ctx *ctx) {
    if (!ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "ctx is null");
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return LEGO_KLT_OK;
}

void *lego_klt_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    // LEGO_KLT_PINNED_WC=1 (experiment): write-combined pinned memory -- faster for the device to read on some hosts,
    // very slow for the CPU to read back
    static const bool wc = getenv("LEGO_KLT_PINNED_WC") && atoi(getenv("LEGO_KLT_PINNED_WC")) != 0;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void lego_klt_free_pinned(void *p) {
    if (p) cudaFreeHost(p);
}

int lego_klt_batch_create(lego_klt_ctx *ctx, int batch, int cols, int rows, size_t step, int n_per_pair,
                          int levels, lego_klt_batch **out) {
    if (out) *out = nullptr;
    int rc = batch_alloc(ctx, batch, cols, rows, step, n_per_pair, levels, out);
    if (rc == LEGO_KLT_OK) {
        (*out)->counted = true;
        ++ctx->live_handles;
    }
    return rc;
}

void lego_klt_batch_destroy(lego_klt_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->maps) warp_maps_destroy(b->maps);
    pyramid_plan_destroy(&b->plan);
    for (int i = 0; i < EV_COUNT; ++i)
        if (b->ev[i]) cudaEventDestroy(b->ev[i]);
    for (int r = 0; r < kRing; ++r)
        for (int i = 0; i < 3; ++i)
            if (b->ring[r][i]) cudaEventDestroy(b->ring[r][i]);
    cudaFree(b->d_pair_count);
    cudaFree(b->d_detect);
    cudaFree(b->d_defer_list);
    cudaFree(b->d_fam_list);
    cudaFree(b->d_templates);
    cudaFree(b->d_feat_flag);
    cudaFree(b->d_scratch);
    if (b->side) cudaStreamDestroy(b->side);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->copy) cudaStreamDestroy(b->copy);
    for (int i = 0; i < kMaxChunks; ++i)
        if (b->ev_chunk[i]) cudaEventDestroy(b->ev_chunk[i]);
    if (b->ev_compute_done) cudaEventDestroy(b->ev_compute_done);
    if (b->d2h) cudaStreamDestroy(b->d2h);
    for (int i = 0; i < kMaxChunks; ++i) {
        if (b->ev_done[i]) cudaEventDestroy(b->ev_done[i]);
        if (b->ev_join_c[i]) cudaEventDestroy(b->ev_join_c[i]);
    }
    cudaFree(b->d_images);
    cudaFree(b->d_tight);
    cudaFree(b->d_io);
    if (b->h_stats) cudaFreeHost(b->h_stats);
    cudaGetLastError();
    lego_klt_ctx *ctx = b->ctx;
    const bool counted = b->counted;
    delete b;
    if (counted) ctx_release_handle(ctx);
}

int lego_klt_batch_upload(lego_klt_batch *b, const uint8_t *imgs1, const uint8_t *imgs2, const float *kp1_xy,
                          const float *kp2_xy) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (!imgs1 || !imgs2) return fail(LEGO_KLT_ERR_BAD_ARG, "image pointer is null");
    const size_t nt = (size_t)b->B * (size_t)b->n_active;
    if (nt && (!kp1_xy || !kp2_xy)) return fail(LEGO_KLT_ERR_BAD_ARG, "keypoint pointer is null");
    CU_TRY(cudaSetDevice(b->ctx->device));
    NvtxRange range("lego_klt H2D");
    cudaStream_t st = b->ctx->stream;
    CU_TRY(cudaEventRecord(b->ev[EV_START], st));
    // copies first, the re-pitch kernels after them (a copy queued behind a kernel of its own stream holds up the copy
    // engine's queue for other streams' uploads)
    if (nt) {
        CU_TRY(cudaMemcpyAsync(b->d_kp1, kp1_xy, nt * sizeof(float2), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(b->d_kp2_init, kp2_xy, nt * sizeof(float2), cudaMemcpyHostToDevice, st));
    }
    CU_TRY(upload_set(b, 0, imgs1, 0, b->B, st));
    CU_TRY(upload_set(b, 1, imgs2, 0, b->B, st));
    CU_TRY(ingest_set(b, 0, 0, b->B, st));
    CU_TRY(ingest_set(b, 1, 0, b->B, st));
    b->uploaded = true;
    b->pyramids_valid = false;
    b->last_chunked = false;
    b->detect_max = 0;   // (corners of the previous images are not handed over to the new ones)
    return LEGO_KLT_OK;
}

int lego_klt_batch_set_feature_counts(lego_klt_batch *b, const int *counts) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (b == b->ctx->single) return fail(LEGO_KLT_ERR_BAD_ARG, "internal batch");
    if (!counts) {
        b->ragged = false;
        return LEGO_KLT_OK;
    }
    unsigned long long total = 0;
    for (int i = 0; i < b->B; ++i) {
        if (counts[i] < 0 || counts[i] > b->n_cap)
            return fail(LEGO_KLT_ERR_BAD_ARG, "counts[%d] = %d is outside [0, %d]", i, counts[i], b->n_cap);
        total += (unsigned long long)counts[i];
    }
    CU_TRY(cudaSetDevice(b->ctx->device));
    if (!b->d_pair_count) CU_TRY(cudaMalloc(&b->d_pair_count, (size_t)b->B * sizeof(int)));
    CU_TRY(cudaStreamSynchronize(b->ctx->stream));  // an earlier copy may still read the host vector
    b->h_pair_count.assign(counts, counts + b->B);
    CU_TRY(cudaMemcpyAsync(b->d_pair_count, b->h_pair_count.data(), (size_t)b->B * sizeof(int), cudaMemcpyHostToDevice,
                           b->ctx->stream));
    b->ragged = true;
    b->n_valid = total;
    return LEGO_KLT_OK;
}

int lego_klt_batch_set_pipeline_chunks(lego_klt_batch *b, int chunks) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (chunks < 0 || chunks > kMaxChunks) return fail(LEGO_KLT_ERR_BAD_ARG, "chunks must be in [0, %d]", kMaxChunks);
    b->pipeline_chunks = chunks;
    return LEGO_KLT_OK;
}

int lego_klt_batch_run(lego_klt_batch *b, const lego_klt_params *params) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    return batch_run(b, params);
}

int lego_klt_batch_download(lego_klt_batch *b, float *kp2_xy, uint8_t *success, lego_klt_stats *stats) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (!b->ran) return fail(LEGO_KLT_ERR_STATE, "lego_klt_batch_download before run");
    const size_t nt = (size_t)b->B * (size_t)b->n_active;
    if (nt && (!kp2_xy || !success)) return fail(LEGO_KLT_ERR_BAD_ARG, "output pointer is null");
    CU_TRY(cudaSetDevice(b->ctx->device));
    NvtxRange range("lego_klt D2H");
    cudaStream_t st = b->ctx->stream;
    if (nt) {
        CU_TRY(cudaMemcpyAsync(kp2_xy, b->d_kp2_out, nt * sizeof(float2), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(success, b->d_success, nt, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaMemcpyAsync(b->h_stats, b->d_stats, kStatCount * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(b->ev[EV_D2H], st));
    CU_TRY(cudaStreamSynchronize(st));
    if (stats) fill_stats(b, stats);
    return LEGO_KLT_OK;
}

int lego_klt_batch_timings(lego_klt_batch *b, int last_n, float *ms_pyramid_avg, float *ms_solver_avg) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (last_n <= 0 || last_n > kRing || last_n > b->timed_runs)
        return fail(LEGO_KLT_ERR_BAD_ARG, "last_n must be in [1, min(%d, lego_klt_batch_run calls so far)]", kRing);
    CU_TRY(cudaSetDevice(b->ctx->device));
    CU_TRY(cudaStreamSynchronize(b->ctx->stream));
    double sp = 0, ss = 0;
    for (long long r = b->timed_runs - last_n; r < b->timed_runs; ++r) {
        float mp = 0.f, msol = 0.f;
        CU_TRY(run_times(b, (int)(r % kRing), &mp, &msol));
        sp += mp;
        ss += msol;
    }
    if (ms_pyramid_avg) *ms_pyramid_avg = (float)(sp / last_n);
    if (ms_solver_avg) *ms_solver_avg = (float)(ss / last_n);
    return LEGO_KLT_OK;
}

int lego_klt_track_batched(lego_klt_batch *b, const lego_klt_params *params, const uint8_t *imgs1,
                           const uint8_t *imgs2, const float *kp1_xy, float *kp2_xy, uint8_t *success,
                           lego_klt_stats *stats) {
    int rc = lego_klt_track_batched_begin(b, params, imgs1, imgs2, kp1_xy, kp2_xy, success);
    if (rc) return rc;
    return lego_klt_track_batched_end(b, stats);
}

int lego_klt_track_batched_end(lego_klt_batch *b, lego_klt_stats *stats) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (!b->in_flight) return fail(LEGO_KLT_ERR_STATE, "lego_klt_track_batched_end without lego_klt_track_batched_begin");
    b->in_flight = false;
    CU_TRY(cudaSetDevice(b->ctx->device));
    CU_TRY(cudaStreamSynchronize(b->ctx->stream));   // (the result copies are joined into the context stream)
    if (stats) fill_stats(b, stats);
    return LEGO_KLT_OK;
}

int lego_klt_track_batched_begin(lego_klt_batch *b, const lego_klt_params *params, const uint8_t *imgs1,
                                 const uint8_t *imgs2, const float *kp1_xy, float *kp2_xy, uint8_t *success) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (b->in_flight) return fail(LEGO_KLT_ERR_STATE, "lego_klt_track_batched_begin: the previous call has not been ended");
    b->detect_max = 0;
    // Small batches: plain upload -> run -> download.
    // Equal chunks: more / smaller chunks shorten the pipeline tail but cost launches and solver efficiency
    // (measured: 10 chunks with a fine tail 6 % slower than 8 equal ones; a smaller LAST chunk of 8 or 16 pairs 5 % slower).
    int bounds[kMaxChunks + 1];
    // Chunk count (measured, 256 pairs x 2000 features, tracks/s end to end): integer source keypoints (freshly
    // detected corners) 4 / 6 / 8 / 12 chunks -> 0.97 / 1.00 / 1.01 / 0.91e8; sub-pixel ones (tracked points fed back)
    // 0.88 / 0.85 / 0.74 / 0.54e8, because a quarter of them take the two-family path, whose second persistent launch
    // needs larger chunks to stay efficient.  One default for both (no look at the caller's data);
    // lego_klt_batch_set_pipeline_chunks overrides it.
    int n_chunks = b->B >= 32 ? 6 : (b->B >= 8 ? 4 : 1);
    if (b->pipeline_chunks > 0) n_chunks = std::min(std::min(b->pipeline_chunks, kMaxChunks), b->B);
    for (int c = 0; c <= n_chunks; ++c) bounds[c] = (int)((long long)b->B * c / n_chunks);
    if (n_chunks == 1) {
        int rc = lego_klt_batch_upload(b, imgs1, imgs2, kp1_xy, kp2_xy);
        if (rc) return rc;
        rc = lego_klt_batch_run(b, params);
        if (rc) return rc;
        const size_t nt1 = (size_t)b->B * (size_t)b->n_active;
        if (nt1 && (!kp2_xy || !success)) return fail(LEGO_KLT_ERR_BAD_ARG, "output pointer is null");
        cudaStream_t st1 = b->ctx->stream;
        if (nt1) {
            CU_TRY(cudaMemcpyAsync(kp2_xy, b->d_kp2_out, nt1 * sizeof(float2), cudaMemcpyDeviceToHost, st1));
            CU_TRY(cudaMemcpyAsync(success, b->d_success, nt1, cudaMemcpyDeviceToHost, st1));
        }
        CU_TRY(cudaMemcpyAsync(b->h_stats, b->d_stats, kStatCount * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st1));
        CU_TRY(cudaEventRecord(b->ev[EV_D2H], st1));
        b->in_flight = true;
        return LEGO_KLT_OK;
    }
    // Large batches: the batch is cut into chunks of pairs; the H2D copy of chunk c+1 (copy stream) overlaps
    // the kernels of chunk c (context stream).  PCIe is the end-to-end bound (933 KB per 1241x376 pair).
    if (!imgs1 || !imgs2) return fail(LEGO_KLT_ERR_BAD_ARG, "image pointer is null");
    const size_t nt = (size_t)b->B * (size_t)b->n_active;
    if (nt && (!kp1_xy || !kp2_xy || !success)) return fail(LEGO_KLT_ERR_BAD_ARG, "keypoint pointer is null");
    int rc = validate_params(params, b->levels);
    if (rc) return rc;
    lego_klt_ctx *ctx = b->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (!b->copy) {
        CU_TRY(cudaStreamCreateWithFlags(&b->copy, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&b->d2h, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxChunks; ++i) {
            CU_TRY(cudaEventCreateWithFlags(&b->ev_chunk[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&b->ev_done[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&b->ev_join_c[i], cudaEventDisableTiming));
        }
        CU_TRY(cudaEventCreateWithFlags(&b->ev_compute_done, cudaEventDisableTiming));
    }
    CU_TRY(cudaEventRecord(b->ev[EV_START], st));
    CU_TRY(cudaStreamWaitEvent(b->copy, b->ev[EV_START], 0));  // the copy stream starts after earlier work
    CU_TRY(cudaStreamWaitEvent(b->d2h, b->ev[EV_START], 0));
    CU_TRY(cudaMemsetAsync(b->d_stats, 0, kStatCount * sizeof(unsigned long long), st));
    CU_TRY(cudaEventRecord(b->ev[EV_H2D], st));
    const int n = b->n_active;
    int last_chunk = -1;
    // All keypoints first, as two copies (per-chunk keypoint copies are 256 KB each: latency-bound gaps on the link),
    // and before any result comes back: kp2_xy is both the initial guess and the output buffer.
    if (nt) {
        CU_TRY(cudaMemcpyAsync(b->d_kp1, kp1_xy, nt * sizeof(float2), cudaMemcpyHostToDevice, b->copy));
        CU_TRY(cudaMemcpyAsync(b->d_kp2_init, kp2_xy, nt * sizeof(float2), cudaMemcpyHostToDevice, b->copy));
    }
    // (a failure inside the loop leaves asynchronous copies from / to the caller's buffers in flight: drain every
    // stream before reporting it)
    auto enqueue_chunks = [&]() -> int {
        for (int c = 0; c < n_chunks; ++c) {
            const int img0 = bounds[c], img1 = bounds[c + 1];
            const int nimg = img1 - img0;
            if (nimg <= 0) continue;
            {
                NvtxRange range("lego_klt H2D chunk");
                CU_TRY(upload_set(b, 0, imgs1, img0, nimg, b->copy));
                CU_TRY(upload_set(b, 1, imgs2, img0, nimg, b->copy));
            }
            CU_TRY(cudaEventRecord(b->ev_chunk[c], b->copy));
            CU_TRY(cudaStreamWaitEvent(st, b->ev_chunk[c], 0));
            CU_TRY(ingest_set(b, 0, img0, nimg, st));
            CU_TRY(ingest_set(b, 1, img0, nimg, st));
            int rc2 = run_range(b, params, img0, nimg, c, nullptr, nullptr, nullptr, b->ev_join_c[c]);
            if (rc2) return rc2;
            last_chunk = c;
            if (n) {  // this chunk's results go home while the next chunk computes
                const size_t off = (size_t)img0 * n, cnt = (size_t)nimg * n;
                CU_TRY(cudaEventRecord(b->ev_done[c], st));
                CU_TRY(cudaStreamWaitEvent(b->d2h, b->ev_done[c], 0));
                if (b->side) CU_TRY(cudaStreamWaitEvent(b->d2h, b->ev_join_c[c], 0));  // (its side-stream work, if any)
                CU_TRY(cudaMemcpyAsync(kp2_xy + 2 * off, b->d_kp2_out + off, cnt * sizeof(float2), cudaMemcpyDeviceToHost, b->d2h));
                CU_TRY(cudaMemcpyAsync(success + off, b->d_success + off, cnt, cudaMemcpyDeviceToHost, b->d2h));
            }
        }
        return LEGO_KLT_OK;
    };
    rc = enqueue_chunks();
    if (rc) {
        const std::string why = g_last_error;
        cudaStreamSynchronize(b->copy);
        cudaStreamSynchronize(st);
        if (b->side) cudaStreamSynchronize(b->side);
        cudaStreamSynchronize(b->d2h);
        cudaGetLastError();
        g_last_error = why;
        return rc;
    }
    // the side stream runs its chunks in order: the last join covers them all (counters, results)
    if (last_chunk >= 0 && b->side) CU_TRY(cudaStreamWaitEvent(st, b->ev_join_c[last_chunk], 0));
    CU_TRY(cudaEventRecord(b->ev[EV_SOLVE], st));
    b->uploaded = true;
    b->pyramids_valid = true;
    b->ran = true;
    b->last_chunked = true;
    b->last_timed = false;
    b->last_params = *params;
    ++b->runs;
    CU_TRY(cudaMemcpyAsync(b->h_stats, b->d_stats, kStatCount * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    // the result copies (d2h stream) are joined into the context stream: lego_klt_track_batched_end waits on one stream,
    // and an event a caller records on it covers the whole call
    CU_TRY(cudaEventRecord(b->ev_compute_done, b->d2h));
    CU_TRY(cudaStreamWaitEvent(st, b->ev_compute_done, 0));
    CU_TRY(cudaEventRecord(b->ev[EV_D2H], st));
    b->in_flight = true;
    return LEGO_KLT_OK;
}

int lego_klt_batch_device_ptrs(lego_klt_batch *b, void **imgs1, void **imgs2, void **kp1_xy, void **kp2_xy_init,
                               void **kp2_xy_out, void **success) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (imgs1) *imgs1 = b->view.lv[0].base[0];
    if (imgs2) *imgs2 = b->view.lv[0].base[1];
    if (kp1_xy) *kp1_xy = b->d_kp1;
    if (kp2_xy_init) *kp2_xy_init = b->d_kp2_init;
    if (kp2_xy_out) *kp2_xy_out = b->d_kp2_out;
    if (success) *success = b->d_success;
    return LEGO_KLT_OK;
}

int lego_klt_track(lego_klt_ctx *ctx, const lego_klt_params *params, const uint8_t *img1, const uint8_t *img2,
                   int cols, int rows, size_t step, const float *kp1_xy, float *kp2_xy, uint8_t *success, int n,
                   lego_klt_stats *stats) {
    if (!ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "ctx is null");
    if (!params) return fail(LEGO_KLT_ERR_BAD_ARG, "params is null");
    if (!img1 || !img2 || cols <= 0 || rows <= 0 || step < (size_t)cols)
        return fail(LEGO_KLT_ERR_BAD_ARG, "bad image arguments");
    if (n < 0 || (n > 0 && (!kp1_xy || !kp2_xy || !success))) return fail(LEGO_KLT_ERR_BAD_ARG, "bad keypoint arguments");
    if (params->levels < 1 || params->levels > kMaxLevels) return fail(LEGO_KLT_ERR_BAD_ARG, "levels out of range");
    int rc = ensure_single(ctx, cols, rows, step, n, params->levels);
    if (rc) return rc;
    lego_klt_batch *b = ctx->single;
    // Stage through pinned memory: a cv::Mat guarantees only (rows-1)*step + cols readable bytes.
    const size_t img_bytes = (size_t)rows * step;
    CU_TRY(cudaSetDevice(ctx->device));
    lego_klt_ctx::Stage *sl = nullptr;
    rc = acquire_stage(ctx, 2 * img_bytes, &sl);
    if (rc) return rc;
    uint8_t *h1 = sl->p, *h2 = h1 + img_bytes;
    const size_t valid = (size_t)(rows - 1) * step + (size_t)cols;
    memcpy(h1, img1, valid);
    memset(h1 + valid, 0, img_bytes - valid);
    memcpy(h2, img2, valid);
    memset(h2 + valid, 0, img_bytes - valid);
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaEventRecord(b->ev[EV_START], st));
    CU_TRY(upload_set(b, 0, h1, 0, 1, st));
    CU_TRY(upload_set(b, 1, h2, 0, 1, st));
    rc = release_stage(ctx, sl);
    if (rc) return rc;
    CU_TRY(ingest_set(b, 0, 0, 1, st));
    CU_TRY(ingest_set(b, 1, 0, 1, st));
    rc = single_upload_keypoints(ctx, b, kp1_xy, kp2_xy, n);
    if (rc) return rc;
    b->uploaded = true;
    b->pyramids_valid = false;
    b->last_chunked = false;
    rc = batch_run(b, params);
    if (rc) return rc;
    return single_download(ctx, b, kp2_xy, success, n, stats);
}

int lego_klt_image_create(lego_klt_ctx *ctx, int cols, int rows, size_t step, int levels, lego_klt_image **out) {
    if (out) *out = nullptr;
    if (!ctx || !out) return fail(LEGO_KLT_ERR_BAD_ARG, "null ctx/out");
    if (cols <= 0 || rows <= 0 || step < (size_t)cols || levels < 1 || levels > kMaxLevels)
        return fail(LEGO_KLT_ERR_BAD_ARG, "bad image shape");
    int lc[kMaxLevels], lr[kMaxLevels];
    if (!pyramid_level_sizes(cols, rows, levels, lc, lr))
        return fail(LEGO_KLT_ERR_UNSUPPORTED, "pyramid level would be empty for %dx%d, %d levels", cols, rows, levels);
    CU_TRY(cudaSetDevice(ctx->device));
    lego_klt_image *im = new (std::nothrow) lego_klt_image();
    if (!im) return fail(LEGO_KLT_ERR_BAD_ARG, "out of host memory");
    im->ctx = ctx;
    im->cols = cols;
    im->rows = rows;
    im->step = step;
    im->levels = levels;
    memset(&im->view, 0, sizeof(im->view));
    im->view.levels = levels;
    im->view.n_images = 1;
    int pitch[kMaxLevels];
    size_t total = 0, off[kMaxLevels];
    for (int l = 0; l < levels; ++l) {
        LevelView &lv = im->view.lv[l];
        lv.cols = lc[l];
        lv.rows = lr[l];
        lv.step = (l == 0) ? (int)step : lc[l];
        pitch[l] = kApronL + (int)align_up((size_t)lv.step, 16) + kApronR;
        lv.pitch = pitch[l];
        lv.slot = (unsigned long long)lr[l] * pitch[l];
        off[l] = total;
        total += align_up((size_t)lv.slot + 2 * kApronL + 64, 256);
    }
    total += 4096;
    auto cleanup = [&](int code) {
        lego_klt_image_destroy(im);
        return code;
    };
    cudaError_t e = cudaMalloc(&im->d_levels, total);
    if (e != cudaSuccess) return cleanup(fail(LEGO_KLT_ERR_CUDA, "cudaMalloc image levels: %s", cudaGetErrorString(e)));
    e = cudaMemsetAsync(im->d_levels, 0, total, ctx->stream);
    if (e != cudaSuccess) return cleanup(fail(LEGO_KLT_ERR_CUDA, "cudaMemset image: %s", cudaGetErrorString(e)));
    for (int l = 0; l < levels; ++l) im->view.lv[l].base[0] = im->view.lv[l].base[1] = im->d_levels + off[l] + kApronL;
    e = cudaMalloc(&im->d_tight, align_up((size_t)rows * step + 256, 256));
    if (e != cudaSuccess) return cleanup(fail(LEGO_KLT_ERR_CUDA, "cudaMalloc landing buffer: %s", cudaGetErrorString(e)));
    if ((e = pyramid_plan_create(cols, rows, levels, pitch, &im->plan)) != cudaSuccess)
        return cleanup(fail(LEGO_KLT_ERR_CUDA, "pyramid plan: %s", cudaGetErrorString(e)));
    if ((e = warp_maps_create(im->view, &im->maps)) != cudaSuccess)
        return cleanup(fail(LEGO_KLT_ERR_CUDA, "TMA descriptor creation failed: %s", cudaGetErrorString(e)));
    im->counted = true;
    ++ctx->live_handles;
    *out = im;
    return LEGO_KLT_OK;
}

void lego_klt_image_destroy(lego_klt_image *im) {
    if (!im) return;
    cudaSetDevice(im->ctx->device);
    cudaStreamSynchronize(im->ctx->stream);
    if (im->maps) warp_maps_destroy(im->maps);
    pyramid_plan_destroy(&im->plan);
    cudaFree(im->d_levels);
    cudaFree(im->d_tight);
    cudaFree(im->d_full);
    cudaGetLastError();
    lego_klt_ctx *ctx = im->ctx;
    const bool counted = im->counted;
    delete im;
    if (counted) ctx_release_handle(ctx);
}

int lego_klt_image_upload(lego_klt_image *im, const uint8_t *data) {
    if (!im || !data) return fail(LEGO_KLT_ERR_BAD_ARG, "null image/data");
    lego_klt_ctx *ctx = im->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    const size_t img_bytes = (size_t)im->rows * im->step;
    lego_klt_ctx::Stage *sl = nullptr;
    int rc = acquire_stage(ctx, img_bytes, &sl);  // (a ring of pinned slots: no need to drain the stream first)
    if (rc) return rc;
    const size_t valid = (size_t)(im->rows - 1) * im->step + (size_t)im->cols;  // what a cv::Mat guarantees
    memcpy(sl->p, data, valid);
    memset(sl->p + valid, 0, img_bytes - valid);
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaMemcpyAsync(im->d_tight, sl->p, img_bytes, cudaMemcpyHostToDevice, st));
    rc = release_stage(ctx, sl);
    if (rc) return rc;
    CU_TRY(launch_ingest(im->d_tight, im->view.lv[0], 0, 0, 1, st));
    CU_TRY(launch_pyramid(im->plan, im->view, 0, 1, st, 1));
    im->valid = true;
    return LEGO_KLT_OK;
}

int lego_klt_half_size(int v) { return (int)lrint((double)v * 0.5); }  // cvRound: round half to even

int lego_klt_image_upload_fullres(lego_klt_image *im, const uint8_t *full, int full_cols, int full_rows,
                                  size_t full_step) {
    if (!im || !full) return fail(LEGO_KLT_ERR_BAD_ARG, "null image/data");
    if (full_cols <= 0 || full_rows <= 0 || full_step < (size_t)full_cols) return fail(LEGO_KLT_ERR_BAD_ARG, "bad frame shape");
    if (im->cols != lego_klt_half_size(full_cols) || im->rows != lego_klt_half_size(full_rows) || im->step != (size_t)im->cols)
        return fail(LEGO_KLT_ERR_BAD_ARG, "handle is %dx%d step %zu; a %dx%d frame halves to %dx%d (continuous)", im->cols,
                    im->rows, im->step, full_cols, full_rows, lego_klt_half_size(full_cols), lego_klt_half_size(full_rows));
    lego_klt_ctx *ctx = im->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)(full_rows - 1) * full_step + (size_t)full_cols;  // what a cv::Mat guarantees
    lego_klt_ctx::Stage *sl = nullptr;
    int rc = acquire_stage(ctx, bytes, &sl);
    if (rc) return rc;
    if (bytes > im->full_bytes) {
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        if (im->d_full) cudaFree(im->d_full);
        im->d_full = nullptr;
        im->full_bytes = 0;
        CU_TRY(cudaMalloc(&im->d_full, bytes));
        im->full_bytes = bytes;
    }
    memcpy(sl->p, full, bytes);
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaMemcpyAsync(im->d_full, sl->p, bytes, cudaMemcpyHostToDevice, st));
    rc = release_stage(ctx, sl);
    if (rc) return rc;
    CU_TRY(launch_half_nearest(im->d_full, full_cols, full_rows, full_step, im->view.lv[0], st));
    CU_TRY(launch_pyramid(im->plan, im->view, 0, 1, st, 1));
    im->valid = true;
    return LEGO_KLT_OK;
}

int lego_klt_downscale_half(lego_klt_ctx *ctx, const uint8_t *full, int full_cols, int full_rows, size_t full_step,
                            uint8_t *out, size_t out_capacity) {
    if (!ctx || !full || !out) return fail(LEGO_KLT_ERR_BAD_ARG, "null argument");
    const int hc = lego_klt_half_size(full_cols), hr = lego_klt_half_size(full_rows);
    if (full_cols <= 0 || full_rows <= 0 || hc <= 0 || hr <= 0 || full_step < (size_t)full_cols)
        return fail(LEGO_KLT_ERR_BAD_ARG, "bad frame shape");
    if ((size_t)hc * hr > out_capacity) return fail(LEGO_KLT_ERR_BAD_ARG, "output buffer too small");
    lego_klt_image *im = nullptr;
    int rc = lego_klt_image_create(ctx, hc, hr, (size_t)hc, 1, &im);
    if (rc) return rc;
    rc = lego_klt_image_upload_fullres(im, full, full_cols, full_rows, full_step);
    if (rc == LEGO_KLT_OK) {
        const LevelView &lv = im->view.lv[0];
        cudaError_t e = cudaMemcpy2DAsync(out, (size_t)hc, lv.base[0], (size_t)lv.pitch, (size_t)hc, (size_t)hr,
                                          cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(LEGO_KLT_ERR_CUDA, "downscale read-back: %s", cudaGetErrorString(e));
    }
    lego_klt_image_destroy(im);
    return rc;
}

int lego_klt_track_images(lego_klt_ctx *ctx, const lego_klt_params *params, const lego_klt_image *img1,
                          const lego_klt_image *img2, const float *kp1_xy, float *kp2_xy, uint8_t *success, int n,
                          lego_klt_stats *stats) {
    if (!ctx || !params || !img1 || !img2) return fail(LEGO_KLT_ERR_BAD_ARG, "null argument");
    if (img1->ctx != ctx || img2->ctx != ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "image belongs to another context");
    if (!img1->valid || !img2->valid) return fail(LEGO_KLT_ERR_STATE, "lego_klt_track_images before lego_klt_image_upload");
    if (img1->cols != img2->cols || img1->rows != img2->rows || img1->step != img2->step || img1->levels != img2->levels)
        return fail(LEGO_KLT_ERR_BAD_ARG, "img1 and img2 must have the same shape, step and levels");
    if (params->levels != img1->levels)
        return fail(LEGO_KLT_ERR_BAD_ARG, "params->levels (%d) != cached pyramid levels (%d)", params->levels, img1->levels);
    if (n < 0 || (n > 0 && (!kp1_xy || !kp2_xy || !success))) return fail(LEGO_KLT_ERR_BAD_ARG, "bad keypoint arguments");
    int rc = ensure_single(ctx, img1->cols, img1->rows, img1->step, n, params->levels);  // keypoint / scratch buffers
    if (rc) return rc;
    lego_klt_batch *b = ctx->single;
    rc = validate_params(params, b->levels);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaSetDevice(ctx->device));
    rc = single_upload_keypoints(ctx, b, kp1_xy, kp2_xy, n);
    if (rc) return rc;
    PyramidView view = img1->view;
    for (int l = 0; l < view.levels; ++l) view.lv[l].base[1] = img2->view.lv[l].base[0];
    CU_TRY(cudaMemsetAsync(b->d_stats, 0, kStatCount * sizeof(unsigned long long), st));
    // timing events only when the caller asks for counters (the reference's signature has none: the shim passes null)
    if (stats) {
        CU_TRY(cudaEventRecord(b->ev[EV_START], st));
        CU_TRY(cudaEventRecord(b->ev[EV_H2D], st));
    }
    cudaEvent_t *ring = stats ? b->ring[b->timed_runs % kRing] : nullptr;
    rc = run_range(b, params, 0, 1, 0, ring, &view, img2->maps);
    if (rc) return rc;
    if (stats) CU_TRY(cudaEventRecord(b->ev[EV_SOLVE], st));
    ++b->runs;
    if (stats) ++b->timed_runs;
    b->last_timed = stats != nullptr;
    b->ran = true;
    b->last_chunked = false;
    return single_download(ctx, b, kp2_xy, success, n, stats);
}

int lego_klt_track_frame(lego_klt_ctx *ctx, const lego_klt_params *params, const lego_klt_image *prev_left,
                         const lego_klt_image *cur_left, const lego_klt_image *cur_right, const float *kp_prev_xy,
                         float *kp_cur_xy, uint8_t *success_temporal, float *kp_right_xy, uint8_t *success_stereo, int n,
                         lego_klt_stats *stats_temporal, lego_klt_stats *stats_stereo) {
    if (!ctx || !params || !prev_left || !cur_left || !cur_right) return fail(LEGO_KLT_ERR_BAD_ARG, "null argument");
    const lego_klt_image *ims[3] = {prev_left, cur_left, cur_right};
    for (const lego_klt_image *im : ims) {
        if (im->ctx != ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "image belongs to another context");
        if (!im->valid) return fail(LEGO_KLT_ERR_STATE, "lego_klt_track_frame before lego_klt_image_upload");
        if (im->cols != prev_left->cols || im->rows != prev_left->rows || im->step != prev_left->step ||
            im->levels != prev_left->levels)
            return fail(LEGO_KLT_ERR_BAD_ARG, "the three images must have the same shape, step and levels");
    }
    if (params->levels != prev_left->levels)
        return fail(LEGO_KLT_ERR_BAD_ARG, "params->levels (%d) != cached pyramid levels (%d)", params->levels, prev_left->levels);
    if (n < 0 || (n > 0 && (!kp_prev_xy || !kp_cur_xy || !success_temporal || !kp_right_xy || !success_stereo)))
        return fail(LEGO_KLT_ERR_BAD_ARG, "bad keypoint arguments");
    int rc = ensure_single(ctx, prev_left->cols, prev_left->rows, prev_left->step, n, params->levels);
    if (rc == LEGO_KLT_OK) rc = ensure_single(ctx, prev_left->cols, prev_left->rows, prev_left->step, n, params->levels, &ctx->single_b);
    if (rc) return rc;
    lego_klt_batch *bt = ctx->single, *bs = ctx->single_b;
    rc = validate_params(params, bt->levels);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaSetDevice(ctx->device));
    NvtxRange range("lego_klt frame: temporal + stereo");
    // temporal track: last left -> current left (Frontend::TrackLastFrame, src/frontend_g2o.cpp:247-256, 453-492)
    rc = single_upload_keypoints(ctx, bt, kp_prev_xy, kp_cur_xy, n);
    if (rc) return rc;
    PyramidView vt = prev_left->view;
    for (int l = 0; l < vt.levels; ++l) vt.lv[l].base[1] = cur_left->view.lv[l].base[0];
    CU_TRY(cudaMemsetAsync(bt->d_stats, 0, kStatCount * sizeof(unsigned long long), st));
    bt->d_slot_valid = nullptr;
    rc = run_range(bt, params, 0, 1, 0, nullptr, &vt, cur_left->maps);
    if (rc) return rc;
    // stereo match of the features the temporal track kept (Frontend::FindFeaturesInRight, :299-308, 495-535: the
    // current frame's left features are the tracked ones; initial guess = the same pixel, :508): chained on the device,
    // the tracked positions never leave HBM between the two solves
    bs->n_active = n;
    bs->d_kp2_init = bs->d_kp1 + n;
    bs->d_success = reinterpret_cast<uint8_t *>(bs->d_kp2_out + n);
    if (n) {
        CU_TRY(cudaMemcpyAsync(bs->d_kp1, bt->d_kp2_out, (size_t)n * sizeof(float2), cudaMemcpyDeviceToDevice, st));
        CU_TRY(cudaMemcpyAsync(bs->d_kp2_init, bt->d_kp2_out, (size_t)n * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    }
    PyramidView vs = cur_left->view;
    for (int l = 0; l < vs.levels; ++l) vs.lv[l].base[1] = cur_right->view.lv[l].base[0];
    CU_TRY(cudaMemsetAsync(bs->d_stats, 0, kStatCount * sizeof(unsigned long long), st));
    lego_klt_params ps = *params;
    ps.has_initial = 1;
    bs->d_slot_valid = bt->d_success;
    rc = run_range(bs, &ps, 0, 1, 0, nullptr, &vs, cur_right->maps);
    bs->d_slot_valid = nullptr;
    if (rc) return rc;
    ++bt->runs;
    ++bs->runs;
    bt->ran = bs->ran = true;
    bt->last_chunked = bs->last_chunked = false;
    bt->last_timed = bs->last_timed = false;
    // both result groups come home behind ONE synchronisation
    const size_t kp_bytes = (size_t)n * sizeof(float2), grp = kIoHeadBytes + kp_bytes + (size_t)n;
    const size_t need = 2 * align_up(grp, 256) + 64;
    if (ctx->pinned_bytes < need) {
        CU_TRY(cudaStreamSynchronize(st));
        rc = ensure_pinned(ctx, need);
        if (rc) return rc;
    }
    uint8_t *h_t = ctx->pinned, *h_s = ctx->pinned + align_up(grp, 256);
    CU_TRY(cudaMemcpyAsync(h_t, bt->d_io + bt->io_out_off, grp, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(h_s, bs->d_io + bs->io_out_off, grp, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (n) {
        memcpy(kp_cur_xy, h_t + kIoHeadBytes, kp_bytes);
        memcpy(success_temporal, h_t + kIoHeadBytes + kp_bytes, (size_t)n);
        memcpy(kp_right_xy, h_s + kIoHeadBytes, kp_bytes);
        memcpy(success_stereo, h_s + kIoHeadBytes + kp_bytes, (size_t)n);
    }
    if (stats_temporal) {
        memcpy(bt->h_stats, h_t, kStatCount * sizeof(unsigned long long));
        fill_stats(bt, stats_temporal);
    }
    if (stats_stereo) {
        memcpy(bs->h_stats, h_s, kStatCount * sizeof(unsigned long long));
        fill_stats(bs, stats_stereo);
    }
    return LEGO_KLT_OK;
}

int lego_klt_build_pyramid(lego_klt_ctx *ctx, const uint8_t *img, int cols, int rows, size_t step, int levels,
                           uint8_t *out, size_t out_capacity, int *level_cols, int *level_rows) {
    if (!ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "ctx is null");
    if (!img || cols <= 0 || rows <= 0 || step < (size_t)cols || levels < 1 || levels > kMaxLevels)
        return fail(LEGO_KLT_ERR_BAD_ARG, "bad image arguments");
    int rc = ensure_single(ctx, cols, rows, step, 0, levels);
    if (rc) return rc;
    lego_klt_batch *b = ctx->single;
    const size_t img_bytes = (size_t)rows * step;
    rc = ensure_pinned(ctx, 2 * img_bytes);
    if (rc) return rc;
    const size_t valid = (size_t)(rows - 1) * step + (size_t)cols;
    memcpy(ctx->pinned, img, valid);
    memset(ctx->pinned + valid, 0, img_bytes - valid);
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU_TRY(upload_set(b, 0, ctx->pinned, 0, 1, st));
    CU_TRY(upload_set(b, 1, ctx->pinned, 0, 1, st));
    CU_TRY(ingest_set(b, 0, 0, 1, st));
    CU_TRY(ingest_set(b, 1, 0, 1, st));
    CU_TRY(launch_pyramid(b->plan, b->view, 0, 1, st));
    size_t off = 0;
    for (int l = 0; l < levels; ++l) {
        const LevelView &lv = b->view.lv[l];
        if (level_cols) level_cols[l] = lv.cols;
        if (level_rows) level_rows[l] = lv.rows;
        if (l == 0) continue;
        const size_t nbytes = (size_t)lv.cols * lv.rows;
        if (!out || off + nbytes > out_capacity) {
            cudaStreamSynchronize(st);  // earlier level copies into `out` are still in flight
            return fail(LEGO_KLT_ERR_BAD_ARG, "output buffer too small");
        }
        CU_TRY(cudaMemcpy2DAsync(out + off, (size_t)lv.cols, lv.base[0], (size_t)lv.pitch, (size_t)lv.cols,
                                 (size_t)lv.rows, cudaMemcpyDeviceToHost, st));
        off += nbytes;
    }
    CU_TRY(cudaStreamSynchronize(st));
    b->uploaded = false;
    return LEGO_KLT_OK;
}

int lego_klt_debug_read_level(lego_klt_ctx *ctx, int level, uint8_t *out, size_t out_capacity, int *pitch,
                              int *apron_left) {
    if (!ctx || !ctx->single) return fail(LEGO_KLT_ERR_STATE, "lego_klt_debug_read_level before lego_klt_build_pyramid");
    lego_klt_batch *b = ctx->single;
    if (level < 0 || level >= b->levels) return fail(LEGO_KLT_ERR_BAD_ARG, "level out of range");
    const LevelView &lv = b->view.lv[level];
    const size_t nbytes = (size_t)lv.rows * lv.pitch;
    if (!out || nbytes > out_capacity) return fail(LEGO_KLT_ERR_BAD_ARG, "output buffer too small");
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaMemcpyAsync(out, lv.base[0] - kApronL, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    if (pitch) *pitch = lv.pitch;
    if (apron_left) *apron_left = kApronL;
    return LEGO_KLT_OK;
}

// ---- triangulation ---------------------------------------------------------------------------------------
static int ensure_tri(lego_klt_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->tri_bytes) return LEGO_KLT_OK;
    if (ctx->d_tri) cudaFree(ctx->d_tri);
    ctx->d_tri = nullptr;
    ctx->tri_bytes = 0;
    CU_TRY(cudaMalloc(&ctx->d_tri, bytes));
    ctx->tri_bytes = bytes;
    return LEGO_KLT_OK;
}

static size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

int lego_klt_triangulate(lego_klt_ctx *ctx, const double *poses34, int n_views, const double *points_xy, int n,
                     double sing_ratio_thr, double *pt_world, uint8_t *ok) {
    if (!ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "ctx is null");
    if (!poses34 || n_views < 2 || n_views > kTriMaxViews) return fail(LEGO_KLT_ERR_BAD_ARG, "n_views must be in [2, %d]", kTriMaxViews);
    if (n < 0 || (n > 0 && (!points_xy || !pt_world || !ok))) return fail(LEGO_KLT_ERR_BAD_ARG, "bad point arguments");
    if (n == 0) return LEGO_KLT_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    const size_t in_b = up256((size_t)n * n_views * 2 * sizeof(double)), out_b = up256((size_t)n * 3 * sizeof(double));
    int rc = ensure_tri(ctx, in_b + out_b + up256((size_t)n));
    if (rc) return rc;
    double *d_in = reinterpret_cast<double *>(ctx->d_tri), *d_out = reinterpret_cast<double *>(ctx->d_tri + in_b);
    uint8_t *d_ok = ctx->d_tri + in_b + out_b;
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaMemcpyAsync(d_in, points_xy, (size_t)n * n_views * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(launch_triangulate(poses34, n_views, d_in, n, sing_ratio_thr, d_out, d_ok, st));
    CU_TRY(cudaMemcpyAsync(pt_world, d_out, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(ok, d_ok, (size_t)n, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return LEGO_KLT_OK;
}

static int tri_stereo_device(lego_klt_ctx *ctx, const lego_camera *left, const lego_camera *right, const float2 *d_kl,
                             const float2 *d_kr, const uint8_t *d_valid, int n, double thr, size_t scratch_off,
                             double *pt_world, uint8_t *ok) {
    const size_t out_b = up256((size_t)n * 3 * sizeof(double));
    double *d_out = reinterpret_cast<double *>(ctx->d_tri + scratch_off);
    uint8_t *d_ok = ctx->d_tri + scratch_off + out_b;
    double poses[24];
    memcpy(poses, left->pose34, sizeof(double) * 12);
    memcpy(poses + 12, right->pose34, sizeof(double) * 12);
    const double cl[4] = {left->fx, left->fy, left->cx, left->cy}, cr[4] = {right->fx, right->fy, right->cx, right->cy};
    cudaStream_t st = ctx->stream;
    CU_TRY(launch_triangulate_stereo(poses, cl, cr, d_kl, d_kr, d_valid, n, thr, d_out, d_ok, st));
    CU_TRY(cudaMemcpyAsync(pt_world, d_out, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(ok, d_ok, (size_t)n, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return LEGO_KLT_OK;
}

int lego_klt_triangulate_stereo(lego_klt_ctx *ctx, const lego_camera *left, const lego_camera *right,
                            const float *kp_left_xy, const float *kp_right_xy, const uint8_t *valid, int n,
                            double sing_ratio_thr, double *pt_world, uint8_t *ok) {
    if (!ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "ctx is null");
    if (!left || !right) return fail(LEGO_KLT_ERR_BAD_ARG, "camera is null");
    if (n < 0 || (n > 0 && (!kp_left_xy || !kp_right_xy || !pt_world || !ok))) return fail(LEGO_KLT_ERR_BAD_ARG, "bad point arguments");
    if (n == 0) return LEGO_KLT_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    const size_t kp_b = up256((size_t)n * sizeof(float2)), v_b = up256((size_t)n);
    int rc = ensure_tri(ctx, 2 * kp_b + v_b + up256((size_t)n * 3 * sizeof(double)) + up256((size_t)n));
    if (rc) return rc;
    float2 *d_kl = reinterpret_cast<float2 *>(ctx->d_tri), *d_kr = reinterpret_cast<float2 *>(ctx->d_tri + kp_b);
    uint8_t *d_valid = valid ? ctx->d_tri + 2 * kp_b : nullptr;
    cudaStream_t st = ctx->stream;
    CU_TRY(cudaMemcpyAsync(d_kl, kp_left_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_kr, kp_right_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, st));
    if (valid) CU_TRY(cudaMemcpyAsync(d_valid, valid, (size_t)n, cudaMemcpyHostToDevice, st));
    return tri_stereo_device(ctx, left, right, d_kl, d_kr, d_valid, n, sing_ratio_thr, 2 * kp_b + v_b, pt_world, ok);
}

int lego_klt_batch_triangulate(lego_klt_batch *b, const lego_camera *left, const lego_camera *right,
                               double sing_ratio_thr, double *pt_world, uint8_t *ok) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (!left || !right) return fail(LEGO_KLT_ERR_BAD_ARG, "camera is null");
    if (!b->ran) return fail(LEGO_KLT_ERR_STATE, "lego_klt_batch_triangulate before lego_klt_batch_run");
    const size_t nt = (size_t)b->B * (size_t)b->n_active;
    if (nt == 0) return LEGO_KLT_OK;
    if (!pt_world || !ok) return fail(LEGO_KLT_ERR_BAD_ARG, "output pointer is null");
    lego_klt_ctx *ctx = b->ctx;
    CU_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_tri(ctx, up256(nt * 3 * sizeof(double)) + up256(nt));
    if (rc) return rc;
    return tri_stereo_device(ctx, left, right, b->d_kp1, b->d_kp2_out, b->d_success, (int)nt, sing_ratio_thr, 0, pt_world, ok);
}

// ---- feature detection -------------------------------------------------------------------
namespace {

struct GfttArgs {
    const uint8_t *mask;
    size_t mask_step;
    const float *exclude_xy;
    int n_exclude;
    float exclude_half;
    int max_corners;
    double quality;
    double min_distance;
};

int gftt_check(const GfttArgs &g, const float *corners, const int *n_corners) {
    if (!corners || !n_corners) return fail(LEGO_KLT_ERR_BAD_ARG, "output pointer is null");
    if (g.max_corners <= 0) return fail(LEGO_KLT_ERR_BAD_ARG, "max_corners must be positive");
    if (!(g.quality > 0.0) || !(g.min_distance >= 0.0)) return fail(LEGO_KLT_ERR_BAD_ARG, "bad quality level / minimum distance");
    if (g.n_exclude < 0 || (g.n_exclude > 0 && !g.exclude_xy)) return fail(LEGO_KLT_ERR_BAD_ARG, "bad exclusion list");
    return LEGO_KLT_OK;
}

// d_img: device image (pitch bytes per row).  If d_img is null the image is uploaded from `host_img` (host_step).
int gftt_run(lego_klt_ctx *ctx, const uint8_t *d_img, int pitch, const uint8_t *host_img, size_t host_step, int cols, int rows,
             const GfttArgs &g, float *corners_xy, float *scores, int *n_corners) {
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t px = (size_t)cols * rows;
    const size_t ws = align_up(gftt_workspace_bytes(cols, rows, (float)g.min_distance), 256);
    const size_t img_b = d_img ? 0 : align_up((size_t)rows * host_step, 256), mask_b = g.mask ? align_up(px, 256) : 0;
    const size_t ex_b = align_up((size_t)g.n_exclude * sizeof(float2) + 8, 256);
    const size_t out_b = align_up((size_t)g.max_corners * (sizeof(float2) + sizeof(float)) + 64, 256);
    const size_t need = ws + img_b + mask_b + ex_b + out_b;
    if (need > ctx->gftt_bytes) {
        CU_TRY(cudaStreamSynchronize(st));
        if (ctx->d_gftt) cudaFree(ctx->d_gftt);
        ctx->d_gftt = nullptr;
        ctx->gftt_bytes = 0;
        CU_TRY(cudaMalloc(&ctx->d_gftt, need));
        ctx->gftt_bytes = need;
    }
    uint8_t *p = ctx->d_gftt + ws;
    uint8_t *d_up = p;
    p += img_b;
    uint8_t *d_mask = g.mask ? p : nullptr;
    p += mask_b;
    float2 *d_ex = reinterpret_cast<float2 *>(p);
    p += ex_b;
    float2 *d_corners = reinterpret_cast<float2 *>(p);
    float *d_scores = reinterpret_cast<float *>(d_corners + g.max_corners);
    int *d_n = reinterpret_cast<int *>(d_scores + g.max_corners);
    // host inputs go through one pinned staging slot (caller memory is usually pageable)
    const size_t stage_b = img_b + mask_b + ex_b;
    if (stage_b > ex_b || g.n_exclude > 0) {
        lego_klt_ctx::Stage *sl = nullptr;
        int rc = acquire_stage(ctx, stage_b, &sl);
        if (rc) return rc;
        uint8_t *h = sl->p;
        if (!d_img) {
            const size_t valid = (size_t)(rows - 1) * host_step + (size_t)cols;   // what a cv::Mat guarantees
            memcpy(h, host_img, valid);
            CU_TRY(cudaMemcpyAsync(d_up, h, valid, cudaMemcpyHostToDevice, st));
            h += img_b;
        }
        if (g.mask) {
            for (int r = 0; r < rows; ++r) memcpy(h + (size_t)r * cols, g.mask + (size_t)r * g.mask_step, (size_t)cols);
            CU_TRY(cudaMemcpyAsync(d_mask, h, px, cudaMemcpyHostToDevice, st));
            h += mask_b;
        }
        if (g.n_exclude > 0) {
            memcpy(h, g.exclude_xy, (size_t)g.n_exclude * sizeof(float2));
            CU_TRY(cudaMemcpyAsync(d_ex, h, (size_t)g.n_exclude * sizeof(float2), cudaMemcpyHostToDevice, st));
        }
        rc = release_stage(ctx, sl);
        if (rc) return rc;
    }
    NvtxRange range("lego_klt detect features");
    CU_TRY(launch_gftt(d_img ? d_img : d_up, cols, rows, d_img ? pitch : (int)host_step, d_mask, d_ex, g.n_exclude,
                       g.exclude_half, g.max_corners, g.quality, (float)g.min_distance, ctx->d_gftt, ws, d_corners, d_scores, d_n,
                       nullptr, st));
    ctx->gftt_cols = cols;
    ctx->gftt_rows = rows;
    int n = 0;
    CU_TRY(cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (n > 0) {
        CU_TRY(cudaMemcpyAsync(corners_xy, d_corners, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, st));
        if (scores) CU_TRY(cudaMemcpyAsync(scores, d_scores, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
    }
    *n_corners = n;
    return LEGO_KLT_OK;
}

}  // namespace

int lego_klt_detect_features(lego_klt_ctx *ctx, const uint8_t *img, int cols, int rows, size_t step, const uint8_t *mask,
                             size_t mask_step, const float *exclude_xy, int n_exclude, float exclude_half, int max_corners,
                             double quality_level, double min_distance, float *corners_xy, float *scores, int *n_corners) {
    if (!ctx) return fail(LEGO_KLT_ERR_BAD_ARG, "ctx is null");
    if (!img || cols < 3 || rows < 3 || step < (size_t)cols) return fail(LEGO_KLT_ERR_BAD_ARG, "bad image arguments");
    if (mask && mask_step < (size_t)cols) return fail(LEGO_KLT_ERR_BAD_ARG, "bad mask step");
    const GfttArgs g{mask, mask_step, exclude_xy, n_exclude, exclude_half, max_corners, quality_level, min_distance};
    int rc = gftt_check(g, corners_xy, n_corners);
    if (rc) return rc;
    return gftt_run(ctx, nullptr, 0, img, step, cols, rows, g, corners_xy, scores, n_corners);
}

int lego_klt_image_detect_features(lego_klt_image *im, const float *exclude_xy, int n_exclude, float exclude_half, int max_corners,
                                   double quality_level, double min_distance, float *corners_xy, float *scores, int *n_corners) {
    if (!im) return fail(LEGO_KLT_ERR_BAD_ARG, "image is null");
    if (!im->valid) return fail(LEGO_KLT_ERR_STATE, "lego_klt_image_detect_features before lego_klt_image_upload");
    if (im->cols < 3 || im->rows < 3) return fail(LEGO_KLT_ERR_BAD_ARG, "image too small");
    const GfttArgs g{nullptr, 0, exclude_xy, n_exclude, exclude_half, max_corners, quality_level, min_distance};
    int rc = gftt_check(g, corners_xy, n_corners);
    if (rc) return rc;
    const LevelView &lv = im->view.lv[0];
    return gftt_run(im->ctx, lv.base[0], lv.pitch, nullptr, 0, im->cols, im->rows, g, corners_xy, scores, n_corners);
}

int lego_klt_batch_detect_features(lego_klt_batch *b, int set, int exclude_source_keypoints, float exclude_half, int max_corners,
                                   double quality_level, double min_distance, float *corners_xy, float *scores, int *n_corners) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (!b->uploaded) return fail(LEGO_KLT_ERR_STATE, "lego_klt_batch_detect_features before the images are uploaded");
    if (set != 0 && set != 1) return fail(LEGO_KLT_ERR_BAD_ARG, "set must be 0 (img1) or 1 (img2)");
    if (b->cols < 3 || b->rows < 3) return fail(LEGO_KLT_ERR_BAD_ARG, "image too small");
    const GfttArgs g{nullptr, 0, nullptr, 0, exclude_half, max_corners, quality_level, min_distance};
    int rc = gftt_check(g, corners_xy, n_corners);
    if (rc) return rc;
    b->detect_max = 0;   // (no detection to hand over until this one has completed)
    // images per pass through the workspace (~25 bytes per pixel and image): as many as ~4 GB hold, at most 256
    const size_t fit = ((size_t)4 << 30) / ((size_t)b->cols * (size_t)b->rows * 25);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)b->B, 256), fit));
    if (const char *c = getenv("LEGO_KLT_DETECT_CHUNK")) chunk = std::max(1, std::min(chunk, atoi(c)));   // (test hook: several passes)
    if (!gftt_batched_supported(b->cols, b->rows, chunk))
        return fail(LEGO_KLT_ERR_UNSUPPORTED, "batched detection: image too large");
    CU_TRY(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const size_t ws = align_up(gftt_batched_workspace_bytes(b->cols, b->rows, chunk), 256);
    const size_t nc = (size_t)b->B * (size_t)max_corners;
    const size_t out_b = align_up(nc * (sizeof(float2) + sizeof(float)) + (size_t)b->B * sizeof(int) + 64, 256);
    if (ws + out_b > b->detect_bytes) {
        CU_TRY(cudaStreamSynchronize(st));
        if (b->d_detect) cudaFree(b->d_detect);
        b->d_detect = nullptr;
        b->detect_bytes = 0;
        CU_TRY(cudaMalloc(&b->d_detect, ws + out_b));
        b->detect_bytes = ws + out_b;
    }
    float2 *d_corners = reinterpret_cast<float2 *>(b->d_detect + ws);
    float *d_scores = reinterpret_cast<float *>(d_corners + nc);
    int *d_n = reinterpret_cast<int *>(d_scores + nc);
    CU_TRY(cudaMemsetAsync(d_corners, 0, out_b, st));   // (slots beyond a pair's corner count read as zeros)
    const LevelView &lv = b->view.lv[0];
    const bool ex = exclude_source_keypoints != 0 && b->n_active > 0;
    NvtxRange range("lego_klt detect features (batch)");
    CU_TRY(launch_gftt_batched(lv.base[set], lv.slot, lv.pitch, b->cols, b->rows, b->B, ex ? b->d_kp1 : nullptr, ex ? b->n_active : 0,
                               ex && b->ragged ? b->d_pair_count : nullptr, exclude_half, max_corners, quality_level, (float)min_distance,
                               b->d_detect, ws, chunk, d_corners, scores ? d_scores : nullptr, d_n, st));
    CU_TRY(cudaMemcpyAsync(corners_xy, d_corners, nc * sizeof(float2), cudaMemcpyDeviceToHost, st));
    if (scores) CU_TRY(cudaMemcpyAsync(scores, d_scores, nc * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(n_corners, d_n, (size_t)b->B * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    b->detect_max = max_corners;
    b->detect_ws = ws;
    b->detect_counts.assign(n_corners, n_corners + b->B);
    return LEGO_KLT_OK;
}

int lego_klt_batch_use_detected_features(lego_klt_batch *b) {
    if (!b) return fail(LEGO_KLT_ERR_BAD_ARG, "batch is null");
    if (b->detect_max <= 0) return fail(LEGO_KLT_ERR_STATE, "lego_klt_batch_use_detected_features before lego_klt_batch_detect_features");
    if (b->detect_max > b->n_active)
        return fail(LEGO_KLT_ERR_BAD_ARG, "the detection asked for up to %d corners per image, the batch holds %d features per pair",
                    b->detect_max, b->n_active);
    if (b == b->ctx->single) return fail(LEGO_KLT_ERR_BAD_ARG, "internal batch");
    CU_TRY(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const float2 *d_corners = reinterpret_cast<const float2 *>(b->d_detect + b->detect_ws);
    // corners [B][detect_max] -> source keypoints and initial guesses [B][n_active] (the same pixel: src/frontend_g2o.cpp:508)
    const size_t w = (size_t)b->detect_max * sizeof(float2);
    CU_TRY(cudaMemcpy2DAsync(b->d_kp1, (size_t)b->n_active * sizeof(float2), d_corners, w, w, b->B, cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaMemcpy2DAsync(b->d_kp2_init, (size_t)b->n_active * sizeof(float2), d_corners, w, w, b->B, cudaMemcpyDeviceToDevice, st));
    if (!b->d_pair_count) CU_TRY(cudaMalloc(&b->d_pair_count, (size_t)b->B * sizeof(int)));
    CU_TRY(cudaStreamSynchronize(st));  // an earlier copy may still read the host vector
    b->h_pair_count = b->detect_counts;
    unsigned long long total = 0;
    for (int v : b->h_pair_count) total += (unsigned long long)v;
    CU_TRY(cudaMemcpyAsync(b->d_pair_count, b->h_pair_count.data(), (size_t)b->B * sizeof(int), cudaMemcpyHostToDevice, st));
    b->ragged = true;
    b->n_valid = total;
    return LEGO_KLT_OK;
}

int lego_klt_debug_read_eig(lego_klt_ctx *ctx, float *out, size_t capacity, int *cols, int *rows) {
    if (!ctx || !ctx->d_gftt || ctx->gftt_cols <= 0) return fail(LEGO_KLT_ERR_STATE, "lego_klt_debug_read_eig before a detection");
    if (!out || capacity < (size_t)ctx->gftt_cols * ctx->gftt_rows) return fail(LEGO_KLT_ERR_BAD_ARG, "output buffer too small");
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(gftt_debug_eig(ctx->d_gftt, ctx->gftt_cols, ctx->gftt_rows, out, ctx->stream));
    if (cols) *cols = ctx->gftt_cols;
    if (rows) *rows = ctx->gftt_rows;
    return LEGO_KLT_OK;
}

// ---- one process, several devices (SURVEY.md 8b "device list", 8e) --------------------------------------------
constexpr int kMultiMaxDepth = 4;
struct lego_klt_multi {
    struct Shard {
        int device = 0, first = 0, count = 0;
        lego_klt_ctx *ctx = nullptr;
        lego_klt_batch *batch = nullptr;
        // dynamic schedule: two small batch objects (two contexts = two stream sets), so that the upload of one block
        // overlaps the kernels and the result copy of the other
        lego_klt_ctx *dctx[kMultiMaxDepth] = {};
        lego_klt_batch *dbatch[kMultiMaxDepth] = {};
        int last_pairs = 0;   // pairs this device tracked in the last call
        int last_blocks = 0;
    };
    std::vector<Shard> shards;
    std::vector<int> counts;   // ragged batch: per-pair feature counts (empty: all n)
    int B = 0, cols = 0, rows = 0, n = 0, levels = 0;
    int block = 0;             // 0: static contiguous blocks; > 0: devices pull blocks of this many pairs
    int depth = 2;             // blocks in flight per device (dynamic schedule)
    size_t step = 0;
};

}  // extern "C"

namespace {

void multi_free_dynamic(lego_klt_multi *m) {
    for (auto &sh : m->shards)
        for (int j = 0; j < kMultiMaxDepth; ++j) {
            if (sh.dbatch[j]) lego_klt_batch_destroy(sh.dbatch[j]);
            if (sh.dctx[j]) lego_klt_destroy(sh.dctx[j]);
            sh.dbatch[j] = nullptr;
            sh.dctx[j] = nullptr;
        }
}

int multi_ensure_dynamic(lego_klt_multi *m) {
    for (auto &sh : m->shards)
        for (int j = 0; j < m->depth; ++j) {
            if (sh.dbatch[j]) continue;
            int rc = lego_klt_create(sh.device, &sh.dctx[j]);
            if (rc == LEGO_KLT_OK)
                rc = lego_klt_batch_create(sh.dctx[j], m->block, m->cols, m->rows, m->step, m->n, m->levels, &sh.dbatch[j]);
            if (rc != LEGO_KLT_OK) {
                const std::string why = g_last_error;
                multi_free_dynamic(m);
                g_last_error = why;
                return rc;
            }
        }
    return LEGO_KLT_OK;
}

// One block of `npairs` <= b->B pairs through a batch object, everything on its context stream: upload, pyramids,
// solver, result copy.  Returns once the work is enqueued; block_end waits.  `counts`: per-pair feature counts of
// these pairs, or null.
int block_begin(lego_klt_batch *b, const lego_klt_params *params, const uint8_t *imgs1, const uint8_t *imgs2,
                const float *kp1_xy, float *kp2_xy, uint8_t *success, int npairs, const int *counts) {
    int rc = validate_params(params, b->levels);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const size_t nt = (size_t)npairs * (size_t)b->n_active;
    CU_TRY(cudaEventRecord(b->ev[EV_START], st));
    CU_TRY(cudaMemsetAsync(b->d_stats, 0, kStatCount * sizeof(unsigned long long), st));
    // every host -> device copy first, the re-pitch kernels after them: a copy queued behind a kernel of its own stream
    // would hold up the copy engine's queue for the other blocks in flight
    if (nt) {
        CU_TRY(cudaMemcpyAsync(b->d_kp1, kp1_xy, nt * sizeof(float2), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(b->d_kp2_init, kp2_xy, nt * sizeof(float2), cudaMemcpyHostToDevice, st));
    }
    b->ragged = counts != nullptr;
    if (counts) {   // (the previous block of this object has been ended: its copy of h_pair_count is done)
        if (!b->d_pair_count) CU_TRY(cudaMalloc(&b->d_pair_count, (size_t)b->B * sizeof(int)));
        b->h_pair_count.assign(counts, counts + npairs);
        b->h_pair_count.resize(b->B, 0);
        CU_TRY(cudaMemcpyAsync(b->d_pair_count, b->h_pair_count.data(), (size_t)b->B * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    CU_TRY(upload_set(b, 0, imgs1, 0, npairs, st));
    CU_TRY(upload_set(b, 1, imgs2, 0, npairs, st));
    CU_TRY(ingest_set(b, 0, 0, npairs, st));
    CU_TRY(ingest_set(b, 1, 0, npairs, st));
    CU_TRY(cudaEventRecord(b->ev[EV_H2D], st));
    rc = run_range(b, params, 0, npairs, 0, nullptr);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(b->ev[EV_SOLVE], st));
    if (nt) {
        CU_TRY(cudaMemcpyAsync(kp2_xy, b->d_kp2_out, nt * sizeof(float2), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(success, b->d_success, nt, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaMemcpyAsync(b->h_stats, b->d_stats, kStatCount * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(b->ev[EV_D2H], st));
    b->uploaded = true;
    b->pyramids_valid = false;   // (only the first npairs pyramids are this block's)
    b->ran = true;
    b->last_chunked = true;
    b->last_timed = false;
    b->last_params = *params;
    ++b->runs;
    return LEGO_KLT_OK;
}

int block_end(lego_klt_batch *b, lego_klt_stats *stats) {
    CU_TRY(cudaSetDevice(b->ctx->device));
    CU_TRY(cudaStreamSynchronize(b->ctx->stream));
    if (stats) fill_stats(b, stats);
    return LEGO_KLT_OK;
}

void add_stats(lego_klt_stats *acc, const lego_klt_stats &s, bool concurrent) {
    acc->n_success += s.n_success;
    acc->n_nan += s.n_nan;
    acc->n_out_of_image += s.n_out_of_image;
    acc->n_slow_path += s.n_slow_path;
    acc->n_deferred += s.n_deferred;
    for (int l = 0; l < LEGO_KLT_MAX_LEVELS; ++l) acc->gn_iters[l] += s.gn_iters[l];
    for (int k = 0; k < 4; ++k) acc->defer_reason[k] += s.defer_reason[k];
    if (concurrent) {   // devices work concurrently: the slowest one
        acc->ms_h2d = std::max(acc->ms_h2d, s.ms_h2d);
        acc->ms_pyramid = std::max(acc->ms_pyramid, s.ms_pyramid);
        acc->ms_solver = std::max(acc->ms_solver, s.ms_solver);
        acc->ms_d2h = std::max(acc->ms_d2h, s.ms_d2h);
    } else {            // blocks of one device, one after the other
        acc->ms_h2d += s.ms_h2d;
        acc->ms_pyramid += s.ms_pyramid;
        acc->ms_solver += s.ms_solver;
        acc->ms_d2h += s.ms_d2h;
    }
}

}  // namespace

extern "C" {

int lego_klt_multi_create(const int *devices, int n_devices, int batch, int cols, int rows, size_t step, int n_per_pair,
                          int levels, lego_klt_multi **out) {
    if (out) *out = nullptr;
    if (!devices || n_devices <= 0 || !out) return fail(LEGO_KLT_ERR_BAD_ARG, "null / empty device list");
    if (batch < n_devices) return fail(LEGO_KLT_ERR_BAD_ARG, "fewer pairs (%d) than devices (%d)", batch, n_devices);
    lego_klt_multi *m = new (std::nothrow) lego_klt_multi();
    if (!m) return fail(LEGO_KLT_ERR_BAD_ARG, "out of host memory");
    m->B = batch;
    m->cols = cols;
    m->rows = rows;
    m->step = step;
    m->n = n_per_pair;
    m->levels = levels;
    if (const char *d = getenv("LEGO_KLT_MULTI_DEPTH")) m->depth = std::min(std::max(atoi(d), 1), kMultiMaxDepth);  // tuning aid
    m->shards.resize(n_devices);
    // contiguous blocks of pairs whose sizes differ by at most one (the partition of lego_slam_b200/sharding.py)
    const int base = batch / n_devices, rem = batch % n_devices;
    for (int i = 0; i < n_devices; ++i) {
        lego_klt_multi::Shard &sh = m->shards[i];
        sh.device = devices[i];
        sh.first = i * base + (i < rem ? i : rem);
        sh.count = base + (i < rem ? 1 : 0);
        sh.last_pairs = 0;
        int rc = lego_klt_create(sh.device, &sh.ctx);
        if (rc == LEGO_KLT_OK) rc = lego_klt_batch_create(sh.ctx, sh.count, cols, rows, step, n_per_pair, levels, &sh.batch);
        if (rc != LEGO_KLT_OK) {
            const std::string why = g_last_error;
            lego_klt_multi_destroy(m);
            g_last_error = why;
            return rc;
        }
    }
    *out = m;
    return LEGO_KLT_OK;
}

void lego_klt_multi_destroy(lego_klt_multi *m) {
    if (!m) return;
    multi_free_dynamic(m);
    for (auto &sh : m->shards) {
        if (sh.batch) lego_klt_batch_destroy(sh.batch);
        if (sh.ctx) lego_klt_destroy(sh.ctx);
    }
    delete m;
}

int lego_klt_multi_shard(const lego_klt_multi *m, int i, int *device, int *first_pair, int *n_pairs) {
    if (!m || i < 0 || i >= (int)m->shards.size()) return fail(LEGO_KLT_ERR_BAD_ARG, "shard index out of range");
    if (device) *device = m->shards[i].device;
    if (first_pair) *first_pair = m->shards[i].first;
    if (n_pairs) *n_pairs = m->shards[i].count;
    return (int)m->shards.size();
}

int lego_klt_multi_set_feature_counts(lego_klt_multi *m, const int *counts) {
    if (!m) return fail(LEGO_KLT_ERR_BAD_ARG, "multi is null");
    for (auto &sh : m->shards) {
        int rc = lego_klt_batch_set_feature_counts(sh.batch, counts ? counts + sh.first : nullptr);
        if (rc) return rc;
    }
    if (counts) m->counts.assign(counts, counts + m->B);   // (validated block by block above)
    else m->counts.clear();
    return LEGO_KLT_OK;
}

int lego_klt_multi_set_schedule(lego_klt_multi *m, int block_pairs) {
    if (!m) return fail(LEGO_KLT_ERR_BAD_ARG, "multi is null");
    if (block_pairs < 0 || block_pairs > m->B) return fail(LEGO_KLT_ERR_BAD_ARG, "block_pairs must be in [0, %d]", m->B);
    if (block_pairs != m->block) multi_free_dynamic(m);   // (the block objects are sized for one block)
    m->block = block_pairs;
    return LEGO_KLT_OK;
}

int lego_klt_multi_last_distribution(const lego_klt_multi *m, int *pairs_per_device, int capacity) {
    if (!m || !pairs_per_device) return fail(LEGO_KLT_ERR_BAD_ARG, "null argument");
    if (capacity < (int)m->shards.size()) return fail(LEGO_KLT_ERR_BAD_ARG, "capacity < number of devices");
    for (size_t i = 0; i < m->shards.size(); ++i) pairs_per_device[i] = m->shards[i].last_pairs;
    return (int)m->shards.size();
}

int lego_klt_multi_track(lego_klt_multi *m, const lego_klt_params *params, const uint8_t *imgs1, const uint8_t *imgs2,
                         const float *kp1_xy, float *kp2_xy, uint8_t *success, lego_klt_stats *stats) {
    if (!m) return fail(LEGO_KLT_ERR_BAD_ARG, "multi is null");
    if (!params || !imgs1 || !imgs2) return fail(LEGO_KLT_ERR_BAD_ARG, "null argument");
    if ((size_t)m->B * (size_t)m->n && (!kp1_xy || !kp2_xy || !success)) return fail(LEGO_KLT_ERR_BAD_ARG, "keypoint pointer is null");
    const size_t img_bytes = (size_t)m->rows * m->step;
    const int S = (int)m->shards.size();
    std::vector<int> rcs(S, LEGO_KLT_OK);
    std::vector<std::string> errs(S);
    std::vector<lego_klt_stats> st(S);
    for (auto &s1 : st) memset(&s1, 0, sizeof(s1));
    std::atomic<int> next_block(0);
    if (m->block > 0) {
        int rc = multi_ensure_dynamic(m);
        if (rc) return rc;
    }
    const int n_blocks = m->block > 0 ? (m->B + m->block - 1) / m->block : 0;
    const int *counts = m->counts.empty() ? nullptr : m->counts.data();
    auto work_static = [&](int i) {
        lego_klt_multi::Shard &sh = m->shards[i];
        const size_t ko = (size_t)sh.first * (size_t)m->n;
        rcs[i] = lego_klt_track_batched(sh.batch, params, imgs1 + (size_t)sh.first * img_bytes,
                                        imgs2 + (size_t)sh.first * img_bytes, kp1_xy ? kp1_xy + 2 * ko : nullptr,
                                        kp2_xy ? kp2_xy + 2 * ko : nullptr, success ? success + ko : nullptr, &st[i]);
        if (rcs[i]) errs[i] = g_last_error;  // (thread-local in the worker)
        sh.last_pairs = rcs[i] ? 0 : sh.count;
    };
    // Dynamic schedule: the devices pull blocks of pairs from one counter, so that a device behind a slower host link
    // (or a busier one) simply takes fewer of them; two blocks in flight per device.  Pairs are independent: which
    // device tracks a block does not change its bytes.
    auto work_dynamic = [&](int i) {
        lego_klt_multi::Shard &sh = m->shards[i];
        bool pending[kMultiMaxDepth] = {};
        sh.last_pairs = 0;
        sh.last_blocks = 0;
        auto finish = [&](int j) {
            lego_klt_stats s1;
            const int rc = block_end(sh.dbatch[j], &s1);
            pending[j] = false;
            if (rc) {
                if (!rcs[i]) { rcs[i] = rc; errs[i] = g_last_error; }
            } else {
                add_stats(&st[i], s1, false);
            }
        };
        for (int turn = 0; !rcs[i]; ++turn) {
            const int j = turn % m->depth;
            if (pending[j]) finish(j);
            if (rcs[i]) break;
            const int k = next_block.fetch_add(1);
            if (k >= n_blocks) break;
            const int first = k * m->block, np = std::min(m->block, m->B - first);
            const size_t ko = (size_t)first * (size_t)m->n;
            const int rc = block_begin(sh.dbatch[j], params, imgs1 + (size_t)first * img_bytes, imgs2 + (size_t)first * img_bytes,
                                       kp1_xy ? kp1_xy + 2 * ko : nullptr, kp2_xy ? kp2_xy + 2 * ko : nullptr,
                                       success ? success + ko : nullptr, np, counts ? counts + first : nullptr);
            if (rc) {
                rcs[i] = rc;
                errs[i] = g_last_error;
                cudaStreamSynchronize(sh.dbatch[j]->ctx->stream);   // nothing of this block stays in flight
                break;
            }
            pending[j] = true;
            sh.last_pairs += np;
            ++sh.last_blocks;
        }
        for (int q = 0; q < m->depth; ++q) {   // (oldest first)
            const int j = (int)((q + (long long)sh.last_blocks) % m->depth);
            if (pending[j]) finish(j);
        }
    };
    // one host thread per device: each drives its own context(s), streams and pinned-copy pipeline; no exchange between
    // them (pairs are independent), results land in disjoint slices of the caller's buffers
    auto work = [&](int i) {
        if (m->block > 0) work_dynamic(i);
        else work_static(i);
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < S; ++i) pool.emplace_back(work, i);
    work(0);
    for (auto &t : pool) t.join();
    for (int i = 0; i < S; ++i)
        if (rcs[i]) return fail(rcs[i], "device %d: %s", m->shards[i].device, errs[i].c_str());
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (int i = 0; i < S; ++i) add_stats(stats, st[i], true);
        if (counts) {
            for (int v : m->counts) stats->n_features += (uint64_t)v;
        } else {
            stats->n_features = (uint64_t)m->B * (uint64_t)m->n;
        }
    }
    return LEGO_KLT_OK;
}

}  // extern "C"
