// resize_u8.cpp -- CPU restatement of OpenCV's cv::resize for CV_8UC1 / INTER_LINEAR (TEST INFRASTRUCTURE ONLY).
//
// Third-party arithmetic (OpenCV >= 3.2, un-pinned by the reference, CMakeLists.txt:18) behind the reference's call
// sites src/algorithm.cpp:147-150.  PINNED bit-exact against Python cv2.resize 4.13 on every size chain of the
// BASELINE configs and on odd shapes (tests/test_oracle_pyramid.py, hashes in tests/golden/golden.json).
// Linked into libklt_oracle.so (the restated path) and into oracle/_ref/libklt_ref*.so, where it stands in for
// cv::resize under the reference's own, unmodified translation unit (oracle/ref_stubs/opencv2/opencv.hpp).
#include "klt_oracle.h"

#include <cmath>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// cv::resize INTER_LINEAR for CV_8UC1 (third party; call sites src/algorithm.cpp:147-150).
// Restated from OpenCV imgproc resize.cpp: fixed-point coefficients scaled by 2^11, horizontal pass
// to int32, vertical pass  (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.
// ------------------------------------------------------------------------------------------------
struct AxisTable {
    std::vector<int> ofs;       // source index of the left/top tap
    std::vector<short> coef;    // 2 per destination index
};

inline short round_coef(float v) { return (short)std::lrintf(v); }  // saturate_cast<short>(float)

AxisTable make_axis_table(int sn, int dn, bool horizontal) {
    AxisTable t;
    t.ofs.resize(dn);
    t.coef.resize(2 * (size_t)dn);
    double inv_scale = (double)dn / sn;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (horizontal) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= sn - 1) { f = 0; s = sn - 1; }
        }
        t.ofs[d] = s;
        t.coef[2 * d] = round_coef((1.f - f) * 2048.f);
        t.coef[2 * d + 1] = round_coef(f * 2048.f);
    }
    return t;
}

inline int clip_index(int v, int n) { return v < 0 ? 0 : (v < n ? v : n - 1); }

int resize_half_impl(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst) {
    // cv::Size(cols * 0.5, rows * 0.5): int * double -> double, truncated by Size_<int>
    int dw = (int)(sw * 0.5), dh = (int)(sh * 0.5);
    if (dw <= 0 || dh <= 0) return LEGO_KLT_ERR_UNSUPPORTED;
    AxisTable tx = make_axis_table(sw, dw, true);
    AxisTable ty = make_axis_table(sh, dh, false);
    std::vector<int> row0(dw), row1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = clip_index(ty.ofs[dy], sh), sy1 = clip_index(ty.ofs[dy] + 1, sh);
        const uint8_t *s0 = src + (size_t)sy0 * sstep, *s1 = src + (size_t)sy1 * sstep;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = tx.ofs[dx];
            int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            int a0 = tx.coef[2 * dx], a1 = tx.coef[2 * dx + 1];
            row0[dx] = s0[sx] * a0 + s0[sx1] * a1;
            row1[dx] = s1[sx] * a0 + s1[sx1] * a1;
        }
        int b0 = ty.coef[2 * dy], b1 = ty.coef[2 * dy + 1];
        uint8_t *drow = dst + (size_t)dy * dw;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            drow[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    return LEGO_KLT_OK;
}

}  // namespace

extern "C" int klt_oracle_resize_half(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst) {
    if (!src || !dst || sw <= 0 || sh <= 0 || sstep < (size_t)sw) return LEGO_KLT_ERR_BAD_ARG;
    return resize_half_impl(src, sw, sh, sstep, dst);
}
