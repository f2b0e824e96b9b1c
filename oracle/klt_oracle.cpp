// klt_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY, see klt_oracle.h for scope and pinning).
//
// Build like the reference builds its own library (CMakeLists.txt:6-7: -std=c++11 -O3, no -march,
// no -ffast-math): ISO mode => no FMA contraction, SSE2 scalar fp32/fp64 only.
//
// Every function cites the reference lines (relative to the upstream tree) it restates.
#include "klt_oracle.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// 8-bit image plane with the reference's (data, cols, rows, step) view.  `px` must be followed by at
// least step+2 readable zero bytes after row rows-1 (deviation note in klt_oracle.h).
// ------------------------------------------------------------------------------------------------
struct Plane {
    const uint8_t *px;
    int cols, rows;
    size_t step;
};

// Owning, zero-padded copy of an image.
struct PaddedImage {
    std::vector<uint8_t> buf;
    Plane view;
    PaddedImage() : view{nullptr, 0, 0, 0} {}
    void assign(const uint8_t *src, int cols, int rows, size_t step) {
        buf.assign((size_t)rows * step + step + 2, 0);
        for (int r = 0; r < rows; ++r) {
            // copy whole rows including the caller's inter-row padding except on the last row,
            // where only `cols` bytes are guaranteed to exist
            size_t nbytes = (r + 1 < rows) ? step : (size_t)cols;
            std::memcpy(&buf[(size_t)r * step], src + (size_t)r * step, nbytes);
        }
        view = Plane{buf.data(), cols, rows, step};
    }
    void alloc_tight(int cols, int rows) {
        buf.assign((size_t)rows * cols + cols + 2, 0);
        view = Plane{buf.data(), cols, rows, (size_t)cols};
    }
};

// ------------------------------------------------------------------------------------------------
// GetPixelValue -- include/legoslam/algorithm.h:40-57.
// Coordinates are clamped only when outside ( <0  or  >= size ); the four taps are addressed flat:
// p[0], p[1], p[step], p[step+1]; all arithmetic is fp32, products and sums in source order.
// ------------------------------------------------------------------------------------------------
inline float sample_bilinear(const Plane &im, float x, float y) {
    if (x < 0) x = 0;
    if (y < 0) y = 0;
    if (x >= im.cols) x = im.cols - 1;
    if (y >= im.rows) y = im.rows - 1;
    const uint8_t *p = im.px + (size_t)((int)y) * im.step + (int)x;
    float xx = x - std::floor(x);
    float yy = y - std::floor(y);
    return (1 - xx) * (1 - yy) * p[0] + xx * (1 - yy) * p[1] + (1 - xx) * yy * p[im.step] +
           xx * yy * p[im.step + 1];
}

// IsPtInImg -- include/legoslam/algorithm.h:60-66 (float point widened to double).
inline bool point_in_image(float px, float py, const Plane &im) {
    double x = px, y = py;
    return !(x < 0 || y < 0 || x >= im.cols || y >= im.rows);
}

// ------------------------------------------------------------------------------------------------
// Eigen 3.3 LDLT<Matrix2d, Lower>::compute() + _solve_impl() specialised to 2x2
// (third party, restated from Eigen/src/Cholesky/LDLT.h; call site src/algorithm.cpp:93).
//   - pivot: index of the largest |diagonal| entry, first one on ties
//   - a zero pivot at k=0 means the whole diagonal is zero: factorisation stops, D = 0
//   - solve uses the pseudo-inverse of D: components with |D_i| <= tol become 0,
//     tol = 1/highest() in 3.3.x
// ------------------------------------------------------------------------------------------------
inline void ldlt2_solve(double h00, double h10, double h11, double b0, double b1, double &x0,
                        double &x1) {
    bool swapped = std::fabs(h11) > std::fabs(h00);
    double a = swapped ? h11 : h00;  // D(0) candidate
    double d = swapped ? h00 : h11;
    double c = h10;                  // the single sub-diagonal entry is untouched by the 2x2 swap
    double l, d1;
    if (a == 0.0) {
        // "The entire diagonal is zero": L keeps the raw entry, D stays {0, 0}.
        l = c;
        d1 = d;
    } else {
        l = c / a;
        d1 = d - l * (a * l);
    }
    double y0 = swapped ? b1 : b0;
    double y1 = swapped ? b0 : b1;
    y1 = y1 - l * y0;  // L^-1 (unit lower)
    const double tol = 1.0 / std::numeric_limits<double>::max();
    y0 = (std::fabs(a) > tol) ? y0 / a : 0.0;
    y1 = (std::fabs(d1) > tol) ? y1 / d1 : 0.0;
    y0 = y0 - l * y1;  // L^-T (unit upper)
    x0 = swapped ? y1 : y0;
    x1 = swapped ? y0 : y1;
}

// ------------------------------------------------------------------------------------------------
// LKOpticalFlowTracker::calcLKOpticalFlow -- src/algorithm.cpp:37-125, one feature range.
// kp1/kp2 are packed {x,y} floats (cv::KeyPoint::pt).  Mixed precision exactly as the reference:
// fp32 sampling and fp32 differences, widened to fp64 for dx,dy,J,H,b,cost,update.
// ------------------------------------------------------------------------------------------------
struct LevelCounters {
    uint64_t iters = 0, nan = 0;
};

void solve_feature_range(const Plane &img1, const Plane &img2, const float *kp1, float *kp2,
                         uint8_t *success, int begin, int end, const lego_klt_params &prm,
                         bool has_initial, LevelCounters &cnt) {
    const int lo = prm.patch_lo, hi = prm.patch_hi;
    const int half_grad_step = 1;             // :41
    const int iterations = prm.max_iters;     // :42
    const bool inverse = prm.inverse != 0;
    for (int i = begin; i < end; ++i) {
        const float kx = kp1[2 * i], ky = kp1[2 * i + 1];  // :44
        double dx = 0, dy = 0;                              // :45
        if (has_initial) {                                  // :47-50 (float subtraction, widened)
            dx = kp2[2 * i] - kx;
            dy = kp2[2 * i + 1] - ky;
        }
        double cost = 0, lastCost = 0;  // :52
        bool succ = true;               // :53
        double H00 = 0, H01 = 0, H10 = 0, H11 = 0, b0 = 0, b1 = 0;  // :55-56
        double J0 = 0, J1 = 0;  // :57 -- declared OUTSIDE the loops: inverse mode reuses it (F4)
        for (int iter = 0; iter < iterations; ++iter) {  // :58
            if (!inverse) { H00 = H01 = H10 = H11 = 0; }  // :59
            b0 = b1 = 0;                                   // :61
            cost = 0;                                      // :62
            ++cnt.iters;
            for (int x = lo; x <= hi; ++x) {       // :63  x outer
                for (int y = lo; y <= hi; ++y) {   // :64  y inner
                    // :65-66  (kx + x) is a float add; + dx promotes to double; the call narrows
                    double error = sample_bilinear(img1, kx + x, ky + y) -
                                   sample_bilinear(img2, kx + x + dx, ky + y + dy);
                    if (!inverse) {  // :68-73
                        J0 = -1.0 * (0.5 * (sample_bilinear(img2, kx + x + dx + half_grad_step, ky + y + dy) -
                                            sample_bilinear(img2, kx + x + dx - half_grad_step, ky + y + dy)));
                        J1 = -1.0 * (0.5 * (sample_bilinear(img2, kx + x + dx, ky + y + dy + half_grad_step) -
                                            sample_bilinear(img2, kx + x + dx, ky + y + dy - half_grad_step)));
                    } else if (iter == 0) {  // :74-80  (only the first pass refreshes J)
                        J0 = -1.0 * (0.5 * (sample_bilinear(img1, kx + x + half_grad_step, ky + y) -
                                            sample_bilinear(img1, kx + x - half_grad_step, ky + y)));
                        J1 = -1.0 * (0.5 * (sample_bilinear(img1, kx + x, ky + y + half_grad_step) -
                                            sample_bilinear(img1, kx + x, ky + y - half_grad_step)));
                    }
                    b0 += -error * J0;  // :83
                    b1 += -error * J1;
                    cost += error * error;  // :84
                    if (!inverse || iter == 0) {  // :85-87
                        H00 += J0 * J0;
                        H01 += J0 * J1;
                        H10 += J1 * J0;
                        H11 += J1 * J1;
                    }
                }
            }
            double u0, u1;  // :92-93
            ldlt2_solve(H00, H10, H11, b0, b1, u0, u1);
            if (std::isnan(u0) || std::isnan(u1) || std::isinf(u0) || std::isinf(u1)) {  // :94-100
                ++cnt.nan;
                succ = false;
                break;
            }
            if (iter > 0 && cost > lastCost) break;  // :102-104
            dx += u0;  // :107-110
            dy += u1;
            lastCost = cost;
            succ = true;
            if (std::sqrt(u0 * u0 + u1 * u1) < prm.eps) break;  // :113-115  update.norm() < 1e-2
        }
        success[i] = succ ? 1 : 0;                 // :119
        float ox = kx + (float)dx, oy = ky + (float)dy;  // :121  Point2f(dx,dy) narrows, fp32 add
        kp2[2 * i] = ox;
        kp2[2 * i + 1] = oy;
        if (!point_in_image(ox, oy, img2)) success[i] = 0;  // :123
    }
}

// LKOpticalFlow1Layer -- src/algorithm.cpp:11-31: contiguous stripes like cv::parallel_for_.
void solve_level(const Plane &img1, const Plane &img2, const float *kp1, float *kp2,
                 uint8_t *success, int n, const lego_klt_params &prm, bool has_initial, int threads,
                 LevelCounters &cnt) {
    if (threads <= 1 || n < 2 * threads) {
        solve_feature_range(img1, img2, kp1, kp2, success, 0, n, prm, has_initial, cnt);
        return;
    }
    std::vector<std::thread> pool;
    std::vector<LevelCounters> local(threads);
    for (int t = 0; t < threads; ++t) {
        int begin = (int)((long long)n * t / threads), end = (int)((long long)n * (t + 1) / threads);
        pool.emplace_back([&, begin, end, t]() {
            solve_feature_range(img1, img2, kp1, kp2, success, begin, end, prm, has_initial, local[t]);
        });
    }
    for (auto &th : pool) th.join();
    for (auto &c : local) { cnt.iters += c.iters; cnt.nan += c.nan; }
}

// Coarse-to-fine driver -- src/algorithm.cpp:158-205, on prebuilt pyramids.
int track_on_pyramids(const lego_klt_params &prm, const std::vector<Plane> &pyr1,
                      const std::vector<Plane> &pyr2, const float *kp1_xy, float *kp2_xy,
                      uint8_t *success, int n, int threads, lego_klt_stats *stats) {
    const int L = prm.levels;
    const double pyramid_scale = 0.5;                       // :136
    const double scale_top = std::ldexp(1.0, -(L - 1));     // :137,161  scales[pyramids-1]
    std::vector<float> k1(2 * (size_t)n), k2(2 * (size_t)n);
    for (int i = 0; i < 2 * n; ++i) {                       // :160-169  Point2f *= double
        k1[i] = (float)(kp1_xy[i] * scale_top);
        k2[i] = (float)(kp2_xy[i] * scale_top);
    }
    std::vector<LevelCounters> cnt(L);
    for (int level = L - 1; level >= 0; --level) {          // :182-202
        bool has_initial = (level == L - 1) ? (prm.has_initial != 0) : true;  // :185-189
        solve_level(pyr1[level], pyr2[level], k1.data(), k2.data(), success, n, prm, has_initial,
                    threads, cnt[level]);
        if (level > 0) {                                    // :192-201
            for (int i = 0; i < n; ++i) {
                k1[2 * i] = (float)(k1[2 * i] / pyramid_scale);
                k1[2 * i + 1] = (float)(k1[2 * i + 1] / pyramid_scale);
                if (success[i]) {
                    k2[2 * i] = (float)(k2[2 * i] / pyramid_scale);
                    k2[2 * i + 1] = (float)(k2[2 * i + 1] / pyramid_scale);
                } else {
                    k2[2 * i] = k1[2 * i];
                    k2[2 * i + 1] = k1[2 * i + 1];
                }
            }
        }
    }
    std::memcpy(kp2_xy, k2.data(), sizeof(float) * 2 * (size_t)n);  // :205
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->n_features = (uint64_t)n;
        for (int i = 0; i < n; ++i) {
            stats->n_success += success[i];
            if (!point_in_image(kp2_xy[2 * i], kp2_xy[2 * i + 1], pyr2[0])) ++stats->n_out_of_image;
        }
        for (int l = 0; l < L; ++l) {
            stats->gn_iters[l] = cnt[l].iters;
            stats->n_nan += cnt[l].nan;
        }
    }
    return LEGO_KLT_OK;
}

bool params_ok(const lego_klt_params *p) {
    return p && p->levels >= 1 && p->levels <= LEGO_KLT_MAX_LEVELS && p->patch_lo <= p->patch_hi &&
           p->patch_lo >= -16 && p->patch_hi <= 16 && p->max_iters >= 0;
}

// Pyramid part of LKOpticalFlow4Layer -- src/algorithm.cpp:140-154.
int build_levels(const uint8_t *img, int cols, int rows, size_t step, int levels,
                 std::vector<PaddedImage> &out) {
    out.resize(levels);
    out[0].assign(img, cols, rows, step);
    for (int l = 1; l < levels; ++l) {
        const Plane &prev = out[l - 1].view;
        int dw = (int)(prev.cols * 0.5), dh = (int)(prev.rows * 0.5);
        if (dw <= 0 || dh <= 0) return LEGO_KLT_ERR_UNSUPPORTED;
        out[l].alloc_tight(dw, dh);
        int rc = klt_oracle_resize_half(prev.px, prev.cols, prev.rows, prev.step, out[l].buf.data());
        if (rc) return rc;
    }
    return LEGO_KLT_OK;
}

}  // namespace

extern "C" {

int klt_oracle_build_pyramid(const uint8_t *img, int cols, int rows, size_t step, int levels,
                             uint8_t *out, size_t out_capacity, int *level_cols, int *level_rows) {
    if (!img || cols <= 0 || rows <= 0 || step < (size_t)cols || levels < 1 ||
        levels > LEGO_KLT_MAX_LEVELS)
        return LEGO_KLT_ERR_BAD_ARG;
    std::vector<PaddedImage> lv;
    int rc = build_levels(img, cols, rows, step, levels, lv);
    if (rc) return rc;
    size_t off = 0;
    for (int l = 0; l < levels; ++l) {
        if (level_cols) level_cols[l] = lv[l].view.cols;
        if (level_rows) level_rows[l] = lv[l].view.rows;
        if (l == 0) continue;
        size_t nbytes = (size_t)lv[l].view.cols * lv[l].view.rows;
        if (!out || off + nbytes > out_capacity) return LEGO_KLT_ERR_BAD_ARG;
        std::memcpy(out + off, lv[l].buf.data(), nbytes);
        off += nbytes;
    }
    return LEGO_KLT_OK;
}

float klt_oracle_get_pixel_value(const uint8_t *data, int cols, int rows, size_t step,
                                 size_t buf_len, float x, float y) {
    std::vector<uint8_t> padded((size_t)rows * step + step + 2, 0);
    std::memcpy(padded.data(), data, buf_len < padded.size() ? buf_len : padded.size());
    Plane im{padded.data(), cols, rows, step};
    return sample_bilinear(im, x, y);
}

void klt_oracle_ldlt2_solve(const double H[4], const double b[2], double x[2]) {
    ldlt2_solve(H[0], H[2], H[3], b[0], b[1], x[0], x[1]);
}

int klt_oracle_track(const lego_klt_params *params, const uint8_t *img1, const uint8_t *img2,
                     int cols, int rows, size_t step, const float *kp1_xy, float *kp2_xy,
                     uint8_t *success, int n, int threads, lego_klt_stats *stats) {
    if (!params_ok(params) || !img1 || !img2 || cols <= 0 || rows <= 0 || step < (size_t)cols ||
        n < 0 || (n > 0 && (!kp1_xy || !kp2_xy || !success)))
        return LEGO_KLT_ERR_BAD_ARG;
    std::vector<PaddedImage> lv1, lv2;
    int rc = build_levels(img1, cols, rows, step, params->levels, lv1);
    if (rc) return rc;
    rc = build_levels(img2, cols, rows, step, params->levels, lv2);
    if (rc) return rc;
    std::vector<Plane> p1, p2;
    for (int l = 0; l < params->levels; ++l) { p1.push_back(lv1[l].view); p2.push_back(lv2[l].view); }
    return track_on_pyramids(*params, p1, p2, kp1_xy, kp2_xy, success, n, threads, stats);
}

int klt_oracle_track_prebuilt(const lego_klt_params *params, const uint8_t *const *pyr1,
                              const uint8_t *const *pyr2, const int *level_cols,
                              const int *level_rows, const size_t *level_step,
                              const float *kp1_xy, float *kp2_xy, uint8_t *success, int n,
                              int threads, lego_klt_stats *stats) {
    if (!params_ok(params) || !pyr1 || !pyr2 || !level_cols || !level_rows || !level_step || n < 0)
        return LEGO_KLT_ERR_BAD_ARG;
    std::vector<Plane> p1, p2;
    for (int l = 0; l < params->levels; ++l) {
        p1.push_back(Plane{pyr1[l], level_cols[l], level_rows[l], level_step[l]});
        p2.push_back(Plane{pyr2[l], level_cols[l], level_rows[l], level_step[l]});
    }
    return track_on_pyramids(*params, p1, p2, kp1_xy, kp2_xy, success, n, threads, stats);
}

}  // extern "C"
