// The integration INTEGRATION.md prescribes for the reference's own tree, compiled for real: the reference's
// declarations of the two KLT entry points are visible WITH their default arguments (copied prototypes of
// /root/reference include/legoslam/algorithm.h:123-136 -- interface only), an OpenCV look-alike provides cv::Mat /
// cv::KeyPoint and the OPENCV_CORE_HPP guard, and this translation unit plays src/algorithm.cpp: it defines
// LEGOSLAM_GPU_DEFINE_ENTRY_POINTS before including the shim.  A caller then uses the default arguments, like
// Frontend::TrackLastFrame...4LayerSelf would if it passed none.
#include <cstdio>
#include <vector>

#include <opencv2/core.hpp>

namespace legoslam {  // include/legoslam/algorithm.h:123-136, as the reference declares them
void LKOpticalFlow1Layer(const cv::Mat &img1, const cv::Mat &img2, const std::vector<cv::KeyPoint> &kp1,
                         std::vector<cv::KeyPoint> &kp2, std::vector<bool> &success, bool inverse = false,
                         bool has_initial = true);
void LKOpticalFlow4Layer(const cv::Mat &img1, const cv::Mat &img2, const std::vector<cv::KeyPoint> &kp1,
                         std::vector<cv::KeyPoint> &kp2, std::vector<bool> &success, bool inverse = false,
                         bool has_initial = true);
}  // namespace legoslam

#define LEGOSLAM_GPU_DEFINE_ENTRY_POINTS
#include "../../include/legoslam_gpu/algorithm_shim.h"

int main() {
    const int cols = 96, rows = 64;
    std::vector<unsigned char> a(cols * rows), b(cols * rows);
    for (int i = 0; i < cols * rows; ++i) {
        a[i] = (unsigned char)((i * 7 + (i / cols) * 13) & 255);
        b[i] = (unsigned char)(((i + 1) * 7 + (i / cols) * 13) & 255);
    }
    cv::Mat m1{a.data(), cols, rows, {(size_t)cols}}, m2{b.data(), cols, rows, {(size_t)cols}};
    std::vector<cv::KeyPoint> kp1(3), kp2(3);
    for (int i = 0; i < 3; ++i) kp1[i] = kp2[i] = cv::KeyPoint(cv::Point2f{20.f + 10 * i, 30.f}, 7);
    std::vector<bool> ok;
    try {
        legoslam::LKOpticalFlow4Layer(m1, m2, kp1, kp2, ok);           // both defaults
        legoslam::LKOpticalFlow1Layer(m1, m2, kp1, kp2, ok, false);    // one default
    } catch (const std::exception &e) {
        // no GPU in the build container: the library reports it, nothing falls back to a CPU implementation
        std::printf("shim(opencv): %s\n", e.what());
        return 3;
    }
    std::printf("shim(opencv): ok %d %d %d, kp2[0] = (%.3f, %.3f), size kept %.0f\n", (int)ok[0], (int)ok[1], (int)ok[2],
                kp2[0].pt.x, kp2[0].pt.y, kp2[0].size);
    return (ok.size() == 3 && kp2[0].size == 7.f) ? 0 : 4;
}
