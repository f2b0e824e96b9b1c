// ref_capi.cpp -- C entry points of oracle/_ref/libklt_ref*.so (TEST INFRASTRUCTURE ONLY).
//
// The library is the reference's OWN hot-path translation unit (/root/reference/src/algorithm.cpp with
// /root/reference/include/legoslam/algorithm.h, both compiled where they lie, unmodified) on top of the stand-in
// headers of oracle/ref_stubs/.  These wrappers only marshal plain buffers into the std::vector<cv::KeyPoint> /
// cv::Mat arguments the way Frontend::TrackLastFrameLKOpticalFlow4LayerSelf does (src/frontend_g2o.cpp:453-492:
// cv::KeyPoint(pt, 7)), and call legoslam::LKOpticalFlow4Layer / LKOpticalFlow1Layer.
#include "legoslam/algorithm.h"

#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#ifndef KLT_REF_HALF_PATCH
#define KLT_REF_HALF_PATCH 3   // src/algorithm.cpp:40
#endif
#ifndef KLT_REF_PYRAMIDS
#define KLT_REF_PYRAMIDS 4     // src/algorithm.cpp:135
#endif

namespace {

// Level-0 Mats alias the caller's images in the reference (src/algorithm.cpp:143-144); here they alias a copy
// with step+2 zero bytes behind the last row (defined out-of-image reads, see the opencv.hpp stand-in).
cv::Mat padded_copy(const uint8_t *src, int cols, int rows, size_t step, std::vector<uint8_t> &store) {
    store.assign((size_t)rows * step + step + 2, 0);
    for (int r = 0; r < rows; ++r)
        std::memcpy(&store[(size_t)r * step], src + (size_t)r * step, (r + 1 < rows) ? step : (size_t)cols);
    return cv::Mat(rows, cols, store.data(), step);
}

int track_one(const uint8_t *img1, const uint8_t *img2, int cols, int rows, size_t step, const float *kp1_xy,
              float *kp2_xy, uint8_t *success, int n, int inverse, int has_initial, int layers) {
    if (!img1 || !img2 || cols <= 0 || rows <= 0 || step < (size_t)cols || n < 0) return -1;
    if (layers != 1 && layers != KLT_REF_PYRAMIDS) return -4;
    std::vector<uint8_t> s1, s2;
    cv::Mat m1 = padded_copy(img1, cols, rows, step, s1), m2 = padded_copy(img2, cols, rows, step, s2);
    std::vector<cv::KeyPoint> kp1((size_t)n), kp2((size_t)n);
    for (int i = 0; i < n; ++i) {
        kp1[i] = cv::KeyPoint(cv::Point2f(kp1_xy[2 * i], kp1_xy[2 * i + 1]), 7);
        kp2[i] = cv::KeyPoint(cv::Point2f(kp2_xy[2 * i], kp2_xy[2 * i + 1]), 7);
    }
    std::vector<bool> ok;
    if (layers == 1)
        legoslam::LKOpticalFlow1Layer(m1, m2, kp1, kp2, ok, inverse != 0, has_initial != 0);
    else
        legoslam::LKOpticalFlow4Layer(m1, m2, kp1, kp2, ok, inverse != 0, has_initial != 0);
    if ((int)kp2.size() != n || (int)ok.size() != n) return -5;
    for (int i = 0; i < n; ++i) {
        kp2_xy[2 * i] = kp2[i].pt.x;
        kp2_xy[2 * i + 1] = kp2[i].pt.y;
        success[i] = ok[i] ? 1 : 0;
    }
    return 0;
}

}  // namespace

extern "C" {

int klt_ref_half_patch(void) { return KLT_REF_HALF_PATCH; }
int klt_ref_pyramids(void) { return KLT_REF_PYRAMIDS; }
// 1 when the reference text was compiled with its own literals untouched, 0 for a parametrised build
int klt_ref_verbatim(void) {
#ifdef KLT_REF_PARAMETRISED
    return 0;
#else
    return 1;
#endif
}

// legoslam::LKOpticalFlow4Layer (layers == klt_ref_pyramids()) or LKOpticalFlow1Layer (layers == 1)
int klt_ref_track(const uint8_t *img1, const uint8_t *img2, int cols, int rows, size_t step, const float *kp1_xy,
                  float *kp2_xy, uint8_t *success, int n, int inverse, int has_initial, int layers) {
    return track_one(img1, img2, cols, rows, step, kp1_xy, kp2_xy, success, n, inverse, has_initial, layers);
}

// B independent pairs (image b at base + b*rows*step, n features each), one pair per worker thread at a time --
// the CPU baseline arm of bench.py.  The reference's own cv::parallel_for_ runs serially inside each call.
int klt_ref_track_pairs(const uint8_t *imgs1, const uint8_t *imgs2, int batch, int cols, int rows, size_t step,
                        const float *kp1_xy, float *kp2_xy, uint8_t *success, int n, int inverse, int has_initial,
                        int layers, int threads) {
    if (batch < 0 || threads < 1) return -1;
    std::atomic<int> next(0), rc(0);
    auto work = [&]() {
        for (;;) {
            int b = next.fetch_add(1);
            if (b >= batch) break;
            size_t io = (size_t)b * rows * step, ko = (size_t)b * n;
            int r = track_one(imgs1 + io, imgs2 + io, cols, rows, step, kp1_xy + 2 * ko, kp2_xy + 2 * ko,
                              success + ko, n, inverse, has_initial, layers);
            if (r) rc.store(r);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    return rc.load();
}

}  // extern "C"
