// Exercises include/legoslam_gpu/algorithm_shim.h the way Frontend::FindFeaturesInRight... does
// (/root/reference src/frontend_g2o.cpp:495-535), with look-alike cv types (OpenCV headers are not in this
// image), and checks the result against the CPU oracle (test infrastructure).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/legoslam_gpu/algorithm_shim.h"
#include "../../oracle/klt_oracle.h"

namespace fakecv {
struct Point2f { float x, y; };
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : pt{0, 0}, size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s) : pt(p), size(s), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct Mat { unsigned char *data; int cols, rows; size_t step; };
}  // namespace fakecv

// look-alikes of Sophus::SE3d / Eigen::Vector3d, as far as legoslam::triangulation touches them
namespace fakeeigen {
struct Mat34 {
    double v[3][4];
    double operator()(int r, int c) const { return v[r][c]; }
};
struct SE3 {
    Mat34 m;
    Mat34 matrix3x4() const { return m; }
};
struct Vec3 {
    double v[3];
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
}  // namespace fakeeigen

// /root/reference test/legoslam_test_triangulation.cpp:5-23 through the shim's legoslam::triangulation
static int triangulation_kat() {
    using namespace fakeeigen;
    const double pw[3] = {30, 20, 10};
    const double ty[3] = {0, -10, 10};
    std::vector<SE3> poses(3);
    std::vector<Vec3> points(3);
    for (int i = 0; i < 3; ++i) {
        // Eigen::Quaterniond(0, 0, 0, 1): w = 0, z = 1 -> rotation by pi about z = diag(-1, -1, 1)
        const double R[3][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, 1}}, t[3] = {0, ty[i], 0};
        double pc[3];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) poses[i].m.v[r][c] = R[r][c];
            poses[i].m.v[r][3] = t[r];
            pc[r] = R[r][0] * pw[0] + R[r][1] * pw[1] + R[r][2] * pw[2] + t[r];
        }
        points[i] = Vec3{{pc[0] / pc[2], pc[1] / pc[2], 1.0}};
    }
    Vec3 est{{0, 0, 0}};
    const bool ok = legoslam::triangulation(poses, points, est);
    std::printf("shim: triangulation -> %d (%.6f %.6f %.6f)\n", (int)ok, est[0], est[1], est[2]);
    if (!ok) return 5;                                                     // EXPECT_TRUE
    for (int k = 0; k < 3; ++k)
        if (std::fabs(est[k] - pw[k]) > 0.01) return 6;                    // EXPECT_NEAR(..., 0.01)
    return 0;
}

int main() {
    if (int rc = triangulation_kat()) return rc;
    const int cols = 620, rows = 188;
    std::vector<unsigned char> left((size_t)cols * rows), right((size_t)cols * rows);
    // smooth pseudo-random texture; right = left shifted by 6 px
    std::vector<float> noise((size_t)(cols + 64) * rows);
    unsigned s = 12345u;
    for (auto &v : noise) { s = s * 1664525u + 1013904223u; v = (float)(s >> 8) / 16777216.f; }
    auto tex = [&](int x, int y) {
        float acc = 0;
        for (int dy = -2; dy <= 2; ++dy)
            for (int dx = -2; dx <= 2; ++dx) {
                int yy = std::min(std::max(y + dy, 0), rows - 1), xx = std::min(std::max(x + dx, 0), cols + 63);
                acc += noise[(size_t)yy * (cols + 64) + xx];
            }
        return (unsigned char)std::min(255.f, acc * (255.f / 25.f));
    };
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) { left[(size_t)y * cols + x] = tex(x + 8, y); right[(size_t)y * cols + x] = tex(x + 14, y); }
    fakecv::Mat m1{left.data(), cols, rows, (size_t)cols}, m2{right.data(), cols, rows, (size_t)cols};
    std::vector<fakecv::KeyPoint> kps_left, kps_right;
    for (int y = 20; y < rows - 20; y += 12)
        for (int x = 20; x < cols - 20; x += 15) {
            kps_left.emplace_back(fakecv::Point2f{(float)x, (float)y}, 7.f);    // frontend_g2o.cpp:500
            kps_right.emplace_back(fakecv::Point2f{(float)x, (float)y}, 7.f);   // :508 same pixel as the guess
        }
    const size_t n = kps_left.size();
    std::vector<float> a(2 * n), b(2 * n);
    for (size_t i = 0; i < n; ++i) { a[2*i] = kps_left[i].pt.x; a[2*i+1] = kps_left[i].pt.y; b[2*i] = kps_right[i].pt.x; b[2*i+1] = kps_right[i].pt.y; }

    std::vector<bool> status;
    legoslam::LKOpticalFlow4Layer(m1, m2, kps_left, kps_right, status, false, true);   // :515

    lego_klt_params p; lego_klt_default_params(&p);
    std::vector<uint8_t> ok(n);
    if (klt_oracle_track(&p, left.data(), right.data(), cols, rows, cols, a.data(), b.data(), ok.data(), (int)n, 1, nullptr)) return 3;
    double maxd = 0; int flagdiff = 0, good = 0;
    for (size_t i = 0; i < n; ++i) {
        maxd = std::max(maxd, (double)std::fabs(kps_right[i].pt.x - b[2*i]));
        maxd = std::max(maxd, (double)std::fabs(kps_right[i].pt.y - b[2*i+1]));
        flagdiff += (status[i] != (ok[i] != 0));
        good += status[i];
        if (kps_right[i].size != 7.f) return 4;  // other KeyPoint fields are preserved
    }
    std::printf("shim: n=%zu tracked=%d max|dpos| vs oracle=%.3g px flag mismatches=%d\n", n, good, maxd, flagdiff);
    return (maxd <= 1e-3 && flagdiff == 0 && status.size() == n) ? 0 : 1;
}
