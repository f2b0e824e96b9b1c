// algorithm_shim.h -- C++ drop-in for the reference's KLT entry points, on top of the C ABI.
//
// Same names, same signatures, same argument meaning as
//   /root/reference include/legoslam/algorithm.h:123-128   legoslam::LKOpticalFlow1Layer
//   /root/reference include/legoslam/algorithm.h:131-136   legoslam::LKOpticalFlow4Layer
//   /root/reference include/legoslam/algorithm.h:11-34     legoslam::triangulation   
// so the two call sites (src/frontend_g2o.cpp:473, :515; identical in src/frontend_lego.cpp:486,528)
// compile unchanged when this header is included instead of the reference's definitions:
//
//     #include <opencv2/core.hpp>                 // cv::Mat, cv::KeyPoint (any OpenCV >= 3.2)
//     #include "legoslam_gpu/algorithm_shim.h"
//     ...
//     legoslam::LKOpticalFlow4Layer(last_frame_->left_img_, current_frame_->left_img_,
//                                   kps_last, kps_current, status, false, true);
//
// Header only; needs liblego_klt.so (include/lego_klt.h) at link time.  The entry points are templates on the matrix /
// keypoint types (any type with data / cols / rows / step and pt.x / pt.y: cv::Mat and cv::KeyPoint bind as they
// are), so that the header compiles and is tested without OpenCV (tests/shim/ uses look-alike PODs).  Inside the
// reference's own tree, where include/legoslam/algorithm.h already declares the two functions, define
// LEGOSLAM_GPU_DEFINE_ENTRY_POINTS in src/algorithm.cpp instead (see the bottom of this file).
//
// Behaviour kept from the reference (src/algorithm.cpp):
//   * kp2 and success are resized to kp1.size() (:17-18, :158); kp2[i].pt is overwritten, every other
//     cv::KeyPoint field of kp2[i] (size = 7 at the call sites, angle, ...) is preserved (:121);
//   * has_initial == false ignores the incoming kp2 positions on the coarsest level (:47-50, :185-189);
//   * inverse == true is the reference's inverse mode including its stale-Jacobian behaviour.
// Differences: errors are reported (std::runtime_error with the library's message) instead of being
// impossible / UB; nothing is printed on NaN (the reference prints "Update is NaN or INF.", :97);
// img1 and img2 must have the same cols, rows and step.  The reference would also accept two images of different
// shapes (src/algorithm.cpp:128-154 builds the two pyramids independently and every GetPixelValue / IsPtInImg uses
// the size of the image it is given), but neither of its call sites does that (last/current and left/right frames of
// one camera rig, src/frontend_g2o.cpp:473,515), and the device layout keeps both images of a pair in one shape:
// such a call throws here rather than computing something else.
#ifndef LEGOSLAM_GPU_ALGORITHM_SHIM_H
#define LEGOSLAM_GPU_ALGORITHM_SHIM_H

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../lego_klt.h"

namespace legoslam {
namespace gpu {

// One context per calling thread (the reference calls the KLT synchronously from the frontend thread).
inline lego_klt_ctx *thread_context(int device = 0) {
    struct Holder {
        lego_klt_ctx *ctx = nullptr;
        ~Holder() { lego_klt_destroy(ctx); }
    };
    static thread_local Holder holder;
    if (!holder.ctx) {
        int rc = lego_klt_create(device, &holder.ctx);
        if (rc != LEGO_KLT_OK) throw std::runtime_error(std::string("lego_klt_create: ") + lego_klt_last_error());
    }
    return holder.ctx;
}

// MatT needs: data (uint8_t*), cols, rows, step (convertible to size_t), channels()/type are the
// caller's business (the reference passes CV_8UC1).  KeyPointT needs: pt.x, pt.y (float).
template <class MatT, class KeyPointT>
void LKOpticalFlowNLayer(const MatT &img1, const MatT &img2, const std::vector<KeyPointT> &kp1,
                         std::vector<KeyPointT> &kp2, std::vector<bool> &success, bool inverse, bool has_initial,
                         int levels) {
    const size_t n = kp1.size();
    kp2.resize(n);       // src/algorithm.cpp:17,158
    success.resize(n);   // :18
    if (img1.cols != img2.cols || img1.rows != img2.rows || (size_t)img1.step != (size_t)img2.step)
        throw std::runtime_error("LKOpticalFlow: img1 and img2 must have the same size and step");
    std::vector<float> a(2 * n), b(2 * n);
    std::vector<uint8_t> ok(n ? n : 1);
    for (size_t i = 0; i < n; ++i) {
        a[2 * i] = kp1[i].pt.x;
        a[2 * i + 1] = kp1[i].pt.y;
        b[2 * i] = kp2[i].pt.x;
        b[2 * i + 1] = kp2[i].pt.y;
    }
    lego_klt_params p;
    lego_klt_default_params(&p);
    p.levels = levels;
    p.inverse = inverse ? 1 : 0;
    p.has_initial = has_initial ? 1 : 0;
    int rc = lego_klt_track(thread_context(), &p, img1.data, img2.data, img1.cols, img1.rows, (size_t)img1.step,
                            a.data(), b.data(), ok.data(), (int)n, nullptr);
    if (rc != LEGO_KLT_OK) throw std::runtime_error(std::string("lego_klt_track: ") + lego_klt_last_error());
    for (size_t i = 0; i < n; ++i) {
        kp2[i].pt.x = b[2 * i];   // :121 -- only pt changes
        kp2[i].pt.y = b[2 * i + 1];
        success[i] = ok[i] != 0;  // :119,123
    }
}

}  // namespace gpu

// ---- the reference's entry points -------------------------------------------------------------------
template <class MatT, class KeyPointT>
void LKOpticalFlow1Layer(const MatT &img1, const MatT &img2, const std::vector<KeyPointT> &kp1,
                         std::vector<KeyPointT> &kp2, std::vector<bool> &success, bool inverse = false,
                         bool has_initial = true) {
    gpu::LKOpticalFlowNLayer(img1, img2, kp1, kp2, success, inverse, has_initial, 1);
}

template <class MatT, class KeyPointT>
void LKOpticalFlow4Layer(const MatT &img1, const MatT &img2, const std::vector<KeyPointT> &kp1,
                         std::vector<KeyPointT> &kp2, std::vector<bool> &success, bool inverse = false,
                         bool has_initial = true) {
    gpu::LKOpticalFlowNLayer(img1, img2, kp1, kp2, success, inverse, has_initial, 4);  // :135 pyramids = 4
}

// ---- legoslam::triangulation (include/legoslam/algorithm.h:11-34) ------------------------------------
// Same name, arguments and return value.  SE3T needs matrix3x4() whose result is indexable as m(r, c)
// (Sophus::SE3d), Vec3T needs operator[] (Eigen::Vector3d): with the reference's typedefs this IS
//     bool triangulation(const std::vector<SE3> &poses, const VecVec3 &points, Vec3 &pt_world, double thr = 1e-3)
// and Frontend::TriangulateNewPoints / BuildInitMap (src/frontend_g2o.cpp:111-155, :310-349) compile unchanged.
// One feature per call costs a launch and two small copies; a frontend that keeps its keypoints on the GPU
// calls lego_klt_batch_triangulate / lego_klt_triangulate_stereo once per frame instead.
template <class SE3T, class Vec3T, class Alloc>
bool triangulation(const std::vector<SE3T> &poses, const std::vector<Vec3T, Alloc> &points, Vec3T &pt_world,
                   double singRatioThr = 1e-3) {
    const size_t nv = poses.size();
    if (nv < 2 || nv > LEGO_TRI_MAX_VIEWS || points.size() < nv)
        throw std::runtime_error("triangulation: 2..8 views with one point each are supported");
    std::vector<double> m(12 * nv), xy(2 * nv);
    for (size_t i = 0; i < nv; ++i) {
        const auto mat = poses[i].matrix3x4();
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) m[12 * i + 4 * r + c] = mat(r, c);
        xy[2 * i] = points[i][0];
        xy[2 * i + 1] = points[i][1];
    }
    double out[3];
    uint8_t ok = 0;
    int rc = lego_klt_triangulate(gpu::thread_context(), m.data(), (int)nv, xy.data(), 1, singRatioThr, out, &ok);
    if (rc != LEGO_KLT_OK) throw std::runtime_error(std::string("lego_klt_triangulate: ") + lego_klt_last_error());
    pt_world[0] = out[0];  // written whatever the verdict, like the reference's out-parameter (:24)
    pt_world[1] = out[1];
    pt_world[2] = out[2];
    return ok != 0;
}

// ---- the reference's own two functions, for the translation unit that used to define them ------------------------
// include/legoslam/algorithm.h:123-136 DECLARES
//     void LKOpticalFlow1Layer(const cv::Mat &, const cv::Mat &, const std::vector<cv::KeyPoint> &,
//                              std::vector<cv::KeyPoint> &, std::vector<bool> &, bool inverse=false, bool has_initial=true);
//     void LKOpticalFlow4Layer(... same ...);
// with their default arguments, and src/algorithm.cpp DEFINES them.  A maintainer keeps the declarations as they are
// and replaces the two definitions in src/algorithm.cpp by
//     #define LEGOSLAM_GPU_DEFINE_ENTRY_POINTS
//     #include "legoslam_gpu/algorithm_shim.h"
// which expands to the two definitions below: ordinary (non-inline, non-template) functions with exactly the declared
// signatures and NO default arguments of their own -- a default argument may be given only once, and the reference's
// header has already given it.  Exactly one translation unit of a program may define the macro (one definition rule).
// Needs the OpenCV core header before this one (OPENCV_CORE_HPP is its include guard).
#if defined(LEGOSLAM_GPU_DEFINE_ENTRY_POINTS) && defined(OPENCV_CORE_HPP)
void LKOpticalFlow1Layer(const cv::Mat &img1, const cv::Mat &img2, const std::vector<cv::KeyPoint> &kp1,
                         std::vector<cv::KeyPoint> &kp2, std::vector<bool> &success, bool inverse, bool has_initial) {
    gpu::LKOpticalFlowNLayer(img1, img2, kp1, kp2, success, inverse, has_initial, 1);
}
void LKOpticalFlow4Layer(const cv::Mat &img1, const cv::Mat &img2, const std::vector<cv::KeyPoint> &kp1,
                         std::vector<cv::KeyPoint> &kp2, std::vector<bool> &success, bool inverse, bool has_initial) {
    gpu::LKOpticalFlowNLayer(img1, img2, kp1, kp2, success, inverse, has_initial, 4);  // src/algorithm.cpp:135
}
#elif defined(LEGOSLAM_GPU_DEFINE_ENTRY_POINTS)
#error "LEGOSLAM_GPU_DEFINE_ENTRY_POINTS needs <opencv2/core.hpp> included first"
#endif

}  // namespace legoslam
#endif  // LEGOSLAM_GPU_ALGORITHM_SHIM_H
