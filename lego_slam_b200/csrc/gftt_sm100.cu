// gftt_sm100.cu -- feature detection of the reference's frontend on the device.
//
// Reference: Frontend::DetectFeatures, src/frontend_g2o.cpp:279-297, with gftt_ = cv::GFTTDetector::create(num_features,
// 0.01, 20) (:16): a mask that is 0 in the rectangle pt +- (10, 10) around every existing left feature, then
// cv::goodFeaturesToTrack(left_img, maxCorners, 0.01, 20, mask, blockSize 3, Shi-Tomasi).  OpenCV is third party; its
// algorithm is restated (oracle/gftt_np.py lists the steps and pins them against Python cv2):
//   gftt_eig_kernel        Sobel 3x3 (fp32, scale 1/3060, BORDER_REFLECT_101) -> products -> 3x3 box sums
//                          (BORDER_REFLECT_101) -> min eigenvalue, fused; also the masked maximum (block reduce + atomicMax)
//   gftt_candidates_kernel threshold at max * qualityLevel, 3x3 non-maximum suppression, mask -> 64-bit keys
//                          (score bits << 32 | flat index), appended with one atomic per warp
//   cub::DeviceRadixSort   keys descending = score descending, ties by higher address first (OpenCV's greaterThanPtr)
//   gftt_select_kernel     the greedy minimum-distance selection, in candidate order, one CTA: every chunk of 1024
//                          candidates is filtered in parallel against the corners accepted so far (a cell grid finer than
//                          minDistance / sqrt 2: at most one corner per cell), the few survivors are taken in order.
// Every fp32 operation is individually rounded and in the oracle's order (this library is built with -fmad=false), so the
// eigenvalue map -- and with it the corner list -- equals the numpy oracle's bit for bit; against OpenCV itself the pin is
// tolerance-aware (its box filter's summation order is not fixed by its published behaviour): tests/test_gftt.py.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>

#include "klt_kernels.h"

namespace legoklt {

namespace {

// BORDER_REFLECT_101 for indices up to n - 1 beyond either end; farther ones (threads of a tile that overhangs the image:
// their results are never used) are clamped so that they stay inside the image.
__device__ __forceinline__ int reflect101(int i, int n) {
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * (n - 1) - i : i;
    return min(max(i, 0), n - 1);
}

constexpr int kEigTilesY = 4;       // 32 x 8 tiles per block of the candidates kernel

// Monotone map float -> unsigned (for atomicMax on floats of either sign).
__device__ __forceinline__ unsigned float_key(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_float(unsigned k) {
    const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// Row-streaming form of the eigenvalue kernel.  A warp owns a strip of kEigStripW = 30 output columns (lanes 1..30; lanes 0
// and 31 are its one-column halo) and walks down kEigChunkRows output rows.  Per pixel row a lane loads the three pixels
// around its column (BORDER_REFLECT_101 on the coordinates: what the Sobel of a position INSIDE the image reads beyond
// the border), forms the row's two 1-D passes, and keeps the last two rows' in registers: the Sobel pair of row y is
// complete when row y + 1 arrives.  The three products go to the neighbouring lanes by shuffle for the horizontal
// 3-sums, which are again kept for three rows for the vertical ones.  Products OUTSIDE the image take the value of their
// mirror position (the box filter's border rule): column -1 := column 1, row -1 := row 1 (and the same at the far
// ends) -- a substitution of whole products, never a Sobel of mirrored pixels (its cross term would change sign).
// Same fp32 operations in the same order as the tile version it replaces (183 instructions per pixel; this one ~80).
constexpr int kEigStripW = 30, kEigChunkRows = 32, kEigWarps = 8;

__global__ void __launch_bounds__(32 * kEigWarps)
gftt_eig_kernel(const uint8_t *__restrict__ img, int cols, int rows, int pitch, const uint8_t *__restrict__ mask, int mask_pitch,
                float *__restrict__ eig, unsigned *__restrict__ max_key, size_t img_stride) {
    __shared__ unsigned block_max;
    {   // batched launch: image blockIdx.z (its own eigenvalue map, mask and maximum)
        const size_t z = blockIdx.z, px = (size_t)cols * rows;
        img += z * img_stride;
        eig += z * px;
        if (mask) mask += z * px;
        max_key += z;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) block_max = 0u;
    __syncthreads();
    const int strip = blockIdx.x * kEigWarps + warp;
    const int c = strip * kEigStripW - 1 + lane;                       // this lane's column (may lie outside the image)
    const int y0 = blockIdx.y * kEigChunkRows, y1 = min(y0 + kEigChunkRows, rows);
    unsigned my_max = 0u;
    if (strip * kEigStripW < cols && y0 < rows) {                       // (uniform per warp)
        const float s = 1.0f / 3060.0f;
        const int cc = min(max(c, 0), cols - 1);                        // halo lanes outside the image: any valid column
        const int cl = reflect101(cc - 1, cols), cr = reflect101(cc + 1, cols);
        const bool first_col = (c == 0), last_col = (c == cols - 1);
        const bool writes = lane >= 1 && lane <= kEigStripW && c >= 0 && c < cols;
        const int ya = max(y0 - 1, 0), yb = min(y1, rows - 1);          // product rows this chunk needs
        float rx_a = 0.f, rx_b = 0.f, sx_a = 0.f, sx_b = 0.f;           // 1-D passes of pixel rows q-2, q-1
        float h2[3] = {0.f, 0.f, 0.f}, h1[3] = {0.f, 0.f, 0.f};         // horizontal 3-sums of product rows y-2, y-1
        auto emit = [&](int yo, const float (&top)[3], const float (&mid)[3], const float (&bot)[3]) {
            const float bxx = __fadd_rn(__fadd_rn(top[0], mid[0]), bot[0]);
            const float bxy = __fadd_rn(__fadd_rn(top[1], mid[1]), bot[1]);
            const float byy = __fadd_rn(__fadd_rn(top[2], mid[2]), bot[2]);
            const float a = __fmul_rn(bxx, 0.5f), b = bxy, cq = __fmul_rn(byy, 0.5f);
            const float amc = __fadd_rn(a, -cq);
            const float e = __fadd_rn(__fadd_rn(a, cq), -__fsqrt_rn(__fadd_rn(__fmul_rn(amc, amc), __fmul_rn(b, b))));
            if (writes) {
                eig[(size_t)yo * cols + c] = e;
                if (!mask || mask[(size_t)yo * mask_pitch + c]) my_max = max(my_max, float_key(e));
            }
        };
        for (int q = ya - 1; q <= yb + 1; ++q) {
            const uint8_t *r = img + (size_t)reflect101(q, rows) * pitch;
            const float pl = (float)__ldg(r + cl), pc = (float)__ldg(r + cc), pr = (float)__ldg(r + cr);
            // row passes of pixel row q: Dx (p[x+1] - p[x-1]) * s;  Dy [1 2 1]
            const float rx_c = __fmul_rn(__fadd_rn(pr, -pl), s);
            const float sx_c = __fadd_rn(__fadd_rn(pl, __fmul_rn(pc, 2.f)), pr);
            if (q >= ya + 1) {
                const int y = q - 1;                                     // the product row completed by pixel row q
                // column passes: Dx [1 2 1] over rows y-1, y, y+1;  Dy (below - above) * s
                const float dx = __fadd_rn(__fadd_rn(rx_a, __fmul_rn(rx_b, 2.f)), rx_c);
                const float dy = __fmul_rn(__fadd_rn(sx_c, -sx_a), s);
                const float p[3] = {__fmul_rn(dx, dx), __fmul_rn(dx, dy), __fmul_rn(dy, dy)};
                float h0[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {                            // (left + centre) + right; mirror at the image's sides
                    const float up = __shfl_up_sync(0xffffffffu, p[k], 1), dn = __shfl_down_sync(0xffffffffu, p[k], 1);
                    const float left = first_col ? dn : up, right = last_col ? up : dn;
                    h0[k] = __fadd_rn(__fadd_rn(left, p[k]), right);
                }
                if (y >= 1 && y - 1 >= y0) {                             // output row y - 1: rows y-2 (row 1 for row 0), y-1, y
                    if (y - 1 == 0) emit(0, h0, h1, h0);
                    else emit(y - 1, h2, h1, h0);
                }
                if (y == rows - 1 && y < y1) emit(y, h1, h0, h1);        // the image's last row: row rows := row rows-2
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    h2[k] = h1[k];
                    h1[k] = h0[k];
                }
            }
            rx_a = rx_b;
            rx_b = rx_c;
            sx_a = sx_b;
            sx_b = sx_c;
        }
    }
    my_max = __reduce_max_sync(0xffffffffu, my_max);   // one shared-memory atomic per warp, one global one per block
    if (lane == 0 && my_max) atomicMax(&block_max, my_max);
    __syncthreads();
    if (threadIdx.x == 0 && block_max) atomicMax(max_key, block_max);
}

// Mask of DetectFeatures (src/frontend_g2o.cpp:280-284): 255, then 0 in pt +- (h, h) for every listed point, both corners
// inclusive; cv::Point2f -> cv::Point rounds half to even (cvRound).
__global__ void gftt_exclusion_kernel(uint8_t *__restrict__ mask, int cols, int rows, const float2 *__restrict__ pts, int n, float half) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const float2 p = pts[i];
    const int x1 = max(__float2int_rn(__fadd_rn(p.x, -half)), 0), y1 = max(__float2int_rn(__fadd_rn(p.y, -half)), 0);
    const int x2 = min(__float2int_rn(__fadd_rn(p.x, half)), cols - 1), y2 = min(__float2int_rn(__fadd_rn(p.y, half)), rows - 1);
    const int w = x2 - x1 + 1, h = y2 - y1 + 1;
    if (w <= 0 || h <= 0) return;
    for (int k = threadIdx.x; k < w * h; k += blockDim.x) mask[(size_t)(y1 + k / w) * cols + x1 + k % w] = 0;
}

// Threshold + 3x3 non-maximum suppression for the pixel of this thread (block = 32 x 8 pixels at tile column blockIdx.x, tile row tile_y).
// Returns the thresholded score; cand = it is a candidate (non-zero, allowed by the mask, equal to the maximum of its 3x3
// neighbourhood; border pixels of the image are never candidates).
__device__ __forceinline__ float nms_candidate(const float *__restrict__ eig, int cols, int rows, const uint8_t *__restrict__ mask,
                                               int mask_pitch, float thr, int tile_y, bool &cand) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = tile_y * 8 + (threadIdx.x >> 5);
    cand = false;
    float v = 0.f;
    if (x >= 1 && y >= 1 && x < cols - 1 && y < rows - 1) {
        v = eig[(size_t)y * cols + x];
        v = v > thr ? v : 0.f;
        if (v != 0.f && (!mask || mask[(size_t)y * mask_pitch + x])) {
            float m = v;   // 3x3 dilation of the thresholded map (all eight neighbours exist here)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const float w = eig[(size_t)(y + dy) * cols + x + dx];
                    m = fmaxf(m, w > thr ? w : 0.f);
                }
            cand = (v == m);
        }
    }
    return v;
}

__global__ void __launch_bounds__(256)
gftt_candidates_kernel(const float *__restrict__ eig, int cols, int rows, const uint8_t *__restrict__ mask, int mask_pitch,
                       const unsigned *__restrict__ max_key, double quality, unsigned long long *__restrict__ keys,
                       unsigned *__restrict__ count) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    // cv::threshold(eig, eig, maxVal * qualityLevel, 0, THRESH_TOZERO) on a CV_32F image compares with (float)thresh
    const float thr = (float)((double)key_float(*max_key) * quality);
    bool cand = false;
    const float v = nms_candidate(eig, cols, rows, mask, mask_pitch, thr, blockIdx.y, cand);
    const unsigned ballot = __ballot_sync(0xffffffffu, cand);
    if (ballot) {
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == __ffs(ballot) - 1) base = atomicAdd(count, (unsigned)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
        if (cand)
            keys[base + __popc(ballot & ((1u << lane) - 1u))] =
                ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(y * cols + x);
    }
}

// Greedy selection in candidate order.  cell_of[] (gw x gh ints, -1 = empty) holds the flat pixel index of the corner
// accepted in a cell of side `cell` <= minDistance / sqrt 2, so a cell never holds two corners and every corner closer
// than minDistance lies within `reach` cells.  Per chunk of 1024 candidates: (A) every thread tests its candidate against
// the corners of EARLIER chunks through the grid (global memory, 1024 lookups in flight); (B) warp 0 walks the survivors
// in order and tests each against the corners accepted in THIS chunk, which it keeps in shared memory (32 comparisons
// per step) -- no dependent global read in the serial part (the first version spent 0.29 ms on 150 corners there).
__global__ void __launch_bounds__(1024)
gftt_select_kernel(const unsigned long long *__restrict__ keys, unsigned n, int cols, int max_corners, float min_dist, int cell,
                   int reach, int gw, int gh, int *__restrict__ cell_of, float2 *__restrict__ corners, float *__restrict__ scores,
                   int *__restrict__ n_out) {
    __shared__ unsigned char alive[1024];
    __shared__ short2 chunk_xy[1024];   // corners accepted in the current chunk
    __shared__ int accepted;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) accepted = 0;
    const float md2 = min_dist * min_dist;
    __syncthreads();
    for (unsigned base = 0; base < n; base += 1024) {
        const unsigned i = base + tid;
        bool live = false;
        if (i < n) {
            live = true;
            if (min_dist >= 1.f) {   // (A) against the corners of earlier chunks
                const unsigned idx = (unsigned)(keys[i] & 0xffffffffull);
                const int x = idx % cols, y = idx / cols, cx = x / cell, cy = y / cell;
                for (int gy = max(cy - reach, 0); gy <= min(cy + reach, gh - 1) && live; ++gy)
                    for (int gx = max(cx - reach, 0); gx <= min(cx + reach, gw - 1); ++gx) {
                        const int q = cell_of[gy * gw + gx];
                        if (q >= 0) {
                            const int dx = q % cols - x, dy = q / cols - y;
                            if ((float)(dx * dx + dy * dy) < md2) live = false;
                        }
                    }
            }
        }
        alive[tid] = live ? 1 : 0;
        __syncthreads();
        if (tid < 32) {   // (B) the survivors in order, by warp 0
            int n_chunk = 0, acc = accepted;
            const unsigned m_total = min(1024u, n - base);
            for (unsigned g = 0; g * 32 < m_total && acc < max_corners; ++g) {
                unsigned m = __ballot_sync(0xffffffffu, g * 32 + lane < m_total && alive[g * 32 + lane]);
                while (m && acc < max_corners) {
                    const unsigned k = g * 32 + (__ffs(m) - 1);
                    m &= m - 1;
                    const unsigned long long key = keys[base + k];
                    const unsigned idx = (unsigned)(key & 0xffffffffull);
                    const int px = idx % cols, py = idx / cols;
                    bool bad = false;
                    if (min_dist >= 1.f)
                        for (int j = lane; j < n_chunk; j += 32) {
                            const int dx = chunk_xy[j].x - px, dy = chunk_xy[j].y - py;
                            bad |= (float)(dx * dx + dy * dy) < md2;
                        }
                    if (__any_sync(0xffffffffu, bad)) continue;
                    if (lane == 0) {
                        chunk_xy[n_chunk] = make_short2((short)px, (short)py);
                        if (min_dist >= 1.f) cell_of[(py / cell) * gw + px / cell] = (int)idx;
                        corners[acc] = make_float2((float)px, (float)py);
                        if (scores) scores[acc] = __uint_as_float((unsigned)(key >> 32));
                    }
                    __syncwarp();
                    ++n_chunk;
                    ++acc;
                }
            }
            if (lane == 0) accepted = acc;
        }
        __syncthreads();   // (also orders warp 0's grid writes before the next chunk's lookups)
        if (accepted >= max_corners) break;
    }
    if (tid == 0) *n_out = accepted;
}

// ---- batched detection: B images of one batch object at once ------------------------------------------------------
// Image b appends its candidate keys (score bits << 32 | flat pixel index, as above) to ITS segment of the key array
// (px slots at b * px) through its own counter; one segmented radix sort orders every segment; one selection CTA per
// image.  No count comes back to the host: the whole detection is asynchronous on the stream.
__global__ void gftt_exclusion_batched_kernel(uint8_t *__restrict__ mask, int cols, int rows, const float2 *__restrict__ pts,
                                              int pts_per_image, const int *__restrict__ counts, float half) {
    const int i = blockIdx.x, b = blockIdx.y;
    if (i >= (counts ? min(counts[b], pts_per_image) : pts_per_image)) return;
    const float2 p = pts[(size_t)b * pts_per_image + i];
    if (!(isfinite(p.x) && isfinite(p.y))) return;
    const int x1 = max(__float2int_rn(__fadd_rn(p.x, -half)), 0), y1 = max(__float2int_rn(__fadd_rn(p.y, -half)), 0);
    const int x2 = min(__float2int_rn(__fadd_rn(p.x, half)), cols - 1), y2 = min(__float2int_rn(__fadd_rn(p.y, half)), rows - 1);
    const int w = x2 - x1 + 1, h = y2 - y1 + 1;
    if (w <= 0 || h <= 0) return;
    uint8_t *m = mask + (size_t)b * cols * rows;
    for (int k = threadIdx.x; k < w * h; k += blockDim.x) m[(size_t)(y1 + k / w) * cols + x1 + k % w] = 0;
}

// Threshold + 3x3 non-maximum suppression of a batch, streaming down rows like gftt_eig_kernel (a warp per 30-column
// strip; the thresholded scores of three rows in registers, neighbours by shuffle).  Candidate keys are collected in a
// per-warp buffer in shared memory and appended to the image's segment with ONE atomic per flush (the tile version's
// atomic per 32 x 8 tile: 469,000 per 256 images; an atomic per warp and row: 3.7 million -- 2.0 ms).
constexpr int kCandBuf = 128;   // keys per warp buffer; flushed when a row could overflow it

__global__ void __launch_bounds__(32 * kEigWarps)
gftt_candidates_batched_kernel(const float *__restrict__ eig_all, int cols, int rows, const uint8_t *__restrict__ mask_all,
                               const unsigned *__restrict__ max_keys, double quality, unsigned long long *__restrict__ keys_all,
                               unsigned *__restrict__ per_image) {
    __shared__ unsigned long long buf[kEigWarps][kCandBuf];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.z;
    const size_t px = (size_t)cols * rows;
    const float *eig = eig_all + (size_t)b * px;
    const uint8_t *mask = mask_all ? mask_all + (size_t)b * px : nullptr;
    unsigned long long *keys = keys_all + (size_t)b * px;
    const int strip = blockIdx.x * kEigWarps + warp;
    const int c = strip * kEigStripW - 1 + lane;
    const int y0 = blockIdx.y * kEigChunkRows, y1 = min(y0 + kEigChunkRows, rows);
    if (strip * kEigStripW >= cols || y0 >= rows) return;   // (uniform per warp; no block barrier below)
    // cv::threshold(eig, eig, maxVal * qualityLevel, 0, THRESH_TOZERO) on a CV_32F image compares with (float)thresh
    const float thr = (float)((double)key_float(max_keys[b]) * quality);
    const bool in_cols = c >= 0 && c < cols;
    // a candidate is an interior pixel of the image (all eight neighbours exist) in an output lane
    const bool may = lane >= 1 && lane <= kEigStripW && c >= 1 && c < cols - 1;
    int n_buf = 0;
    auto flush = [&]() {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(per_image + b, (unsigned)n_buf);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < n_buf; i += 32) keys[base + i] = buf[warp][i];
        __syncwarp();
        n_buf = 0;
    };
    // row maxima over (left, centre, right) of the thresholded scores of rows y-1, y and the centre value of row y
    float m_a = 0.f, m_b = 0.f, v_b = 0.f;
    for (int q = max(y0 - 1, 0); q <= min(y1, rows - 1); ++q) {
        float v = 0.f;
        if (in_cols) {
            v = eig[(size_t)q * cols + c];
            v = v > thr ? v : 0.f;
        }
        const float up = __shfl_up_sync(0xffffffffu, v, 1), dn = __shfl_down_sync(0xffffffffu, v, 1);
        const float m_c = fmaxf(fmaxf(up, v), dn);
        const int y = q - 1;                                   // the row completed by row q
        if (y >= y0 && y >= 1 && y < rows - 1) {               // (y < y1 holds: q <= y1)
            const bool cand = may && v_b != 0.f && (!mask || mask[(size_t)y * cols + c]) && v_b == fmaxf(fmaxf(m_a, m_b), m_c);
            const unsigned ballot = __ballot_sync(0xffffffffu, cand);
            if (ballot) {
                if (n_buf + 32 > kCandBuf) flush();
                if (cand)
                    buf[warp][n_buf + __popc(ballot & ((1u << lane) - 1u))] =
                        ((unsigned long long)__float_as_uint(v_b) << 32) | (unsigned)(y * cols + c);
                n_buf += __popc(ballot);
                __syncwarp();
            }
        }
        m_a = m_b;
        m_b = m_c;
        v_b = v;
    }
    if (n_buf) flush();
}

// segment b of the key array: [b * px, b * px + per_image[b])
__global__ void gftt_segment_offsets_kernel(const unsigned *__restrict__ per_image, int B, int px, int *__restrict__ begin,
                                            int *__restrict__ end) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        begin[b] = b * px;
        end[b] = b * px + (int)per_image[b];
    }
}

// The greedy selection of gftt_select_kernel, one CTA per image.
__global__ void __launch_bounds__(1024)
gftt_select_batched_kernel(const unsigned long long *__restrict__ keys_all, size_t px_per_image,
                           const unsigned *__restrict__ per_image, int cols, int max_corners, float min_dist, int cell, int reach,
                           int gw, int gh, int *__restrict__ cell_all, float2 *__restrict__ corners_all,
                           float *__restrict__ scores_all, int *__restrict__ n_out) {
    __shared__ unsigned char alive[1024];
    __shared__ short2 chunk_xy[1024];
    __shared__ int accepted;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const unsigned long long *keys = keys_all + (size_t)b * px_per_image;
    const unsigned n = per_image[b];
    int *cell_of = cell_all + (size_t)b * gw * gh;
    float2 *corners = corners_all + (size_t)b * max_corners;
    float *scores = scores_all ? scores_all + (size_t)b * max_corners : nullptr;
    const unsigned idx_mask = 0xffffffffu;
    if (tid == 0) accepted = 0;
    const float md2 = min_dist * min_dist;
    __syncthreads();
    for (unsigned base = 0; base < n; base += 1024) {
        const unsigned i = base + tid;
        bool live = false;
        if (i < n) {
            live = true;
            if (min_dist >= 1.f) {
                const unsigned idx = (unsigned)keys[i] & idx_mask;
                const int x = idx % cols, y = idx / cols, cx = x / cell, cy = y / cell;
                for (int gy = max(cy - reach, 0); gy <= min(cy + reach, gh - 1) && live; ++gy)
                    for (int gx = max(cx - reach, 0); gx <= min(cx + reach, gw - 1); ++gx) {
                        const int q = cell_of[gy * gw + gx];
                        if (q >= 0) {
                            const int dx = q % cols - x, dy = q / cols - y;
                            if ((float)(dx * dx + dy * dy) < md2) live = false;
                        }
                    }
            }
        }
        alive[tid] = live ? 1 : 0;
        __syncthreads();
        if (tid < 32) {
            int n_chunk = 0, acc = accepted;
            const unsigned m_total = min(1024u, n - base);
            for (unsigned g = 0; g * 32 < m_total && acc < max_corners; ++g) {
                unsigned m = __ballot_sync(0xffffffffu, g * 32 + lane < m_total && alive[g * 32 + lane]);
                while (m && acc < max_corners) {
                    const unsigned k = g * 32 + (__ffs(m) - 1);
                    m &= m - 1;
                    const unsigned long long key = keys[base + k];
                    const unsigned idx = (unsigned)key & idx_mask;
                    const int px = idx % cols, py = idx / cols;
                    bool bad = false;
                    if (min_dist >= 1.f)
                        for (int j = lane; j < n_chunk; j += 32) {
                            const int dx = chunk_xy[j].x - px, dy = chunk_xy[j].y - py;
                            bad |= (float)(dx * dx + dy * dy) < md2;
                        }
                    if (__any_sync(0xffffffffu, bad)) continue;
                    if (lane == 0) {
                        chunk_xy[n_chunk] = make_short2((short)px, (short)py);
                        if (min_dist >= 1.f) cell_of[(py / cell) * gw + px / cell] = (int)idx;
                        corners[acc] = make_float2((float)px, (float)py);
                        if (scores) scores[acc] = __uint_as_float((unsigned)(key >> 32));
                    }
                    __syncwarp();
                    ++n_chunk;
                    ++acc;
                }
            }
            if (lane == 0) accepted = acc;
        }
        __syncthreads();
        if (accepted >= max_corners) break;
    }
    if (tid == 0) n_out[b] = accepted;
}

static dim3 eig_grid(int cols, int rows, int images) {
    const int strips = (cols + kEigStripW - 1) / kEigStripW;
    return dim3((strips + kEigWarps - 1) / kEigWarps, (rows + kEigChunkRows - 1) / kEigChunkRows, images);
}

}  // namespace

// ---- batched entry ------------------------------------------------------------------------------------------------
// Workspace for `chunk` images at a time: eigenvalue maps, masks, keys and sorted keys (one slot per pixel: the
// suppression keeps ties, so plateaus of equal scores are all candidates), cell grids, scalars, cub's temporary storage.
size_t gftt_batched_workspace_bytes(int cols, int rows, int chunk) {
    const size_t px = (size_t)cols * rows, cap = (size_t)chunk * px;
    size_t sort_tmp = 0;
    cub::DeviceSegmentedRadixSort::SortKeysDescending(nullptr, sort_tmp, (const unsigned long long *)nullptr,
                                                      (unsigned long long *)nullptr, (int)cap, chunk, (const int *)nullptr,
                                                      (const int *)nullptr);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return up(cap * 4) + up(cap) + 2 * up(cap * 8) + up(cap * 4 + 4096) + up((size_t)(4 * chunk + 2) * 4) + sort_tmp + 4096;
}

bool gftt_batched_supported(int cols, int rows, int chunk) {   // (segment offsets are ints)
    return (size_t)cols * rows * (size_t)chunk < ((size_t)1 << 31);
}

cudaError_t launch_gftt_batched(const uint8_t *d_imgs, size_t img_stride, int pitch, int cols, int rows, int n_images,
                                const float2 *d_exclude, int exclude_per_image, const int *d_exclude_counts, float exclude_half,
                                int max_corners, double quality, float min_distance, uint8_t *ws, size_t ws_bytes, int chunk,
                                float2 *d_corners, float *d_scores_or_null, int *d_n_out, cudaStream_t stream) {
    const size_t px = (size_t)cols * rows, cap = (size_t)chunk * px;
    if (!gftt_batched_supported(cols, rows, chunk) || ws_bytes < gftt_batched_workspace_bytes(cols, rows, chunk)) return cudaErrorInvalidValue;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    int cell = 1, reach = 0, gw = 1, gh = 1;
    if (min_distance >= 1.f) {
        cell = (int)floorf(min_distance / 1.41421356f);
        if (cell < 1) cell = 1;
        reach = (int)ceilf(min_distance / (float)cell);
        gw = (cols + cell - 1) / cell;
        gh = (rows + cell - 1) / cell;
    }
    size_t off = 0;
    float *eig = reinterpret_cast<float *>(ws + off);
    off += up(cap * 4);
    uint8_t *mask = ws + off;
    off += up(cap);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + off);
    off += up(cap * 8);
    unsigned long long *sorted = reinterpret_cast<unsigned long long *>(ws + off);
    off += up(cap * 8);
    int *cell_of = reinterpret_cast<int *>(ws + off);
    off += up(cap * 4 + 4096);   // (worst case: one cell per pixel)
    unsigned *scalars = reinterpret_cast<unsigned *>(ws + off);   // max key / count / segment begin / segment end per image
    off += up((size_t)(4 * chunk + 2) * 4);
    uint8_t *sort_tmp = ws + off;
    size_t sort_tmp_bytes = ws_bytes - off;
    unsigned *max_keys = scalars, *per_image = max_keys + chunk;
    int *seg_begin = reinterpret_cast<int *>(per_image + chunk), *seg_end = seg_begin + chunk;
    const bool masked = d_exclude && exclude_per_image > 0;

    cudaError_t e;
    for (int b0 = 0; b0 < n_images; b0 += chunk) {
        const int nb = n_images - b0 < chunk ? n_images - b0 : chunk;
        if ((e = cudaMemsetAsync(scalars, 0, (size_t)(2 * chunk) * 4, stream)) != cudaSuccess) return e;
        if (masked) {
            if ((e = cudaMemsetAsync(mask, 255, (size_t)nb * px, stream)) != cudaSuccess) return e;
            gftt_exclusion_batched_kernel<<<dim3(exclude_per_image, nb), 128, 0, stream>>>(
                mask, cols, rows, d_exclude + (size_t)b0 * exclude_per_image, exclude_per_image,
                d_exclude_counts ? d_exclude_counts + b0 : nullptr, exclude_half);
            note_launch();
        }
        gftt_eig_kernel<<<eig_grid(cols, rows, nb), 32 * kEigWarps, 0, stream>>>(
            d_imgs + (size_t)b0 * img_stride, cols, rows, pitch, masked ? mask : nullptr, cols, eig, max_keys, img_stride);
        note_launch();
        gftt_candidates_batched_kernel<<<eig_grid(cols, rows, nb), 32 * kEigWarps, 0, stream>>>(
            eig, cols, rows, masked ? mask : nullptr, max_keys, quality, keys, per_image);
        note_launch();
        gftt_segment_offsets_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(per_image, nb, (int)px, seg_begin, seg_end);
        note_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        e = cub::DeviceSegmentedRadixSort::SortKeysDescending(sort_tmp, sort_tmp_bytes, keys, sorted, (int)((size_t)nb * px), nb,
                                                              seg_begin, seg_end, 0, 64, stream);
        if (e != cudaSuccess) return e;
        note_launch(2);
        if (min_distance >= 1.f)
            if ((e = cudaMemsetAsync(cell_of, 0xff, (size_t)nb * gw * gh * sizeof(int), stream)) != cudaSuccess) return e;
        gftt_select_batched_kernel<<<nb, 1024, 0, stream>>>(sorted, px, per_image, cols, max_corners, min_distance, cell, reach, gw,
                                                            gh, cell_of, d_corners + (size_t)b0 * max_corners,
                                                            d_scores_or_null ? d_scores_or_null + (size_t)b0 * max_corners : nullptr,
                                                            d_n_out + b0);
        note_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

namespace {
}  // namespace

size_t gftt_workspace_bytes(int cols, int rows, float min_distance) {
    const size_t px = (size_t)cols * rows;
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortKeysDescending(nullptr, sort_tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                             (int)px);
    (void)min_distance;
    return px * 4 /*eig*/ + px /*mask*/ + 2 * px * 8 /*keys, sorted keys*/ + px * 4 /*cell grid, worst case*/ + sort_tmp + 4096;
}

cudaError_t launch_gftt(const uint8_t *d_img, int cols, int rows, int pitch, const uint8_t *d_mask_in, const float2 *d_exclude,
                        int n_exclude, float exclude_half, int max_corners, double quality, float min_distance, uint8_t *ws,
                        size_t ws_bytes, float2 *d_corners, float *d_scores_or_null, int *d_n_out, int *h_n_candidates,
                        cudaStream_t stream) {
    const size_t px = (size_t)cols * rows;
    if (ws_bytes < gftt_workspace_bytes(cols, rows, min_distance)) return cudaErrorInvalidValue;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    float *eig = reinterpret_cast<float *>(ws + off);
    off += up(px * 4);
    uint8_t *mask = ws + off;
    off += up(px);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + off);
    off += up(px * 8);
    unsigned long long *sorted = reinterpret_cast<unsigned long long *>(ws + off);
    off += up(px * 8);
    int *cell_of = reinterpret_cast<int *>(ws + off);
    off += up(px * 4);
    unsigned *scalars = reinterpret_cast<unsigned *>(ws + off);   // [0] max key, [1] candidate count
    off += 256;
    uint8_t *sort_tmp = ws + off;
    size_t sort_tmp_bytes = ws_bytes - off;

    cudaError_t e;
    if ((e = cudaMemsetAsync(scalars, 0, 8, stream)) != cudaSuccess) return e;
    const uint8_t *mask_used = nullptr;
    if (d_mask_in || n_exclude > 0) {
        if (d_mask_in) e = cudaMemcpyAsync(mask, d_mask_in, px, cudaMemcpyDeviceToDevice, stream);
        else e = cudaMemsetAsync(mask, 255, px, stream);
        if (e != cudaSuccess) return e;
        if (n_exclude > 0) {
            gftt_exclusion_kernel<<<n_exclude, 128, 0, stream>>>(mask, cols, rows, d_exclude, n_exclude, exclude_half);
            note_launch();
        }
        mask_used = mask;
    }
    gftt_eig_kernel<<<eig_grid(cols, rows, 1), 32 * kEigWarps, 0, stream>>>(d_img, cols, rows, pitch, mask_used, cols, eig, scalars, 0);
    note_launch();
    gftt_candidates_kernel<<<dim3((cols + 31) / 32, (rows + 7) / 8), 256, 0, stream>>>(eig, cols, rows, mask_used, cols, scalars,
                                                                                      quality, keys, scalars + 1);
    note_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    unsigned n_cand = 0;   // the sort needs the count on the host: one 4-byte read-back
    if ((e = cudaMemcpyAsync(&n_cand, scalars + 1, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    if (h_n_candidates) *h_n_candidates = (int)n_cand;
    if (n_cand > 0) {
        e = cub::DeviceRadixSort::SortKeysDescending(sort_tmp, sort_tmp_bytes, keys, sorted, (int)n_cand, 0, 64, stream);
        if (e != cudaSuccess) return e;
        note_launch(3);   // (cub: histogram, scan, onesweep passes)
    }
    int cell = 1, reach = 0, gw = 1, gh = 1;
    if (min_distance >= 1.f) {
        cell = (int)floorf(min_distance / 1.41421356f);
        if (cell < 1) cell = 1;
        reach = (int)ceilf(min_distance / (float)cell);
        gw = (cols + cell - 1) / cell;
        gh = (rows + cell - 1) / cell;
        if ((e = cudaMemsetAsync(cell_of, 0xff, (size_t)gw * gh * sizeof(int), stream)) != cudaSuccess) return e;
    }
    gftt_select_kernel<<<1, 1024, 0, stream>>>(sorted, n_cand, cols, max_corners, min_distance, cell, reach, gw, gh, cell_of,
                                               d_corners, d_scores_or_null, d_n_out);
    note_launch();
    return cudaGetLastError();
}

cudaError_t gftt_debug_eig(const uint8_t *ws, int cols, int rows, float *h_eig, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyAsync(h_eig, ws, (size_t)cols * rows * 4, cudaMemcpyDeviceToHost, stream);
    return e != cudaSuccess ? e : cudaStreamSynchronize(stream);
}

}  // namespace legoklt
