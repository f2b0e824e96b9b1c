// Look-alike of <opencv2/core.hpp> for tests/shim/shim_opencv_test.cpp: the include guard macro the shim keys on and the
// members of cv::Mat / cv::KeyPoint the reference's KLT entry points touch.  OpenCV headers are not in this image.
#ifndef OPENCV_CORE_HPP
#define OPENCV_CORE_HPP
#include <cstddef>
namespace cv {
struct Point2f {
    float x, y;
};
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt{0, 0}, size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s) : pt(p), size(s), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct MatStep {
    size_t v;
    operator size_t() const { return v; }
};
struct Mat {
    unsigned char *data;
    int cols, rows;
    MatStep step;
};
}  // namespace cv
#endif
