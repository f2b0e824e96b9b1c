// Stand-in for <opencv2/opencv.hpp> (TEST INFRASTRUCTURE ONLY -- oracle/_ref build, see oracle/build_ref.py).
//
// The reference's hot-path translation unit, /root/reference/src/algorithm.cpp, is compiled UNMODIFIED, where it
// lies, against this header instead of OpenCV (absent from the image).  Only what that file and
// include/legoslam/algorithm.h touch is provided, with OpenCV's arithmetic semantics (third party, restated from
// opencv2/core/types.hpp and mat.hpp):
//   cv::Point_<T>      operator+  : Point_<T>(saturate_cast<T>(a.x + b.x), ...)     -> fp32 add for Point2f
//                      operator*= (Point_<T>&, double) : a.x = saturate_cast<T>(a.x * b)  -> (float)((double)x * b)
//                      operator/= (Point_<T>&, double) : a.x = saturate_cast<T>(a.x / b)
//   cv::Size_<int>(int, int)      : double arguments truncate (src/algorithm.cpp:148,150)
//   cv::Mat            data / cols / rows / step (size_t like MatStep::operator size_t), header copies share data
//   cv::parallel_for_  contiguous sub-ranges of independent features; run here as ONE range on the calling thread
//                      (the reference's std::vector<bool> writes race across stripes)
//   cv::resize         CV_8UC1 INTER_LINEAR, the restatement in oracle/resize_u8.cpp that is pinned bit-exact
//                      against Python cv2.resize (tests/test_oracle_pyramid.py)
// Deviation: Mat buffers created here carry step+2 zero bytes after the last row, so that GetPixelValue's reads
// below the last row (algorithm.h:48-55, undefined behaviour in the reference) are defined and read 0.
#ifndef LEGO_REF_STUB_OPENCV_HPP
#define LEGO_REF_STUB_OPENCV_HPP

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

extern "C" int klt_oracle_resize_half(const unsigned char *src, int sw, int sh, size_t sstep, unsigned char *dst);

namespace cv {

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <typename T> static inline T saturate_cast(float v) { return (T)v; }

template <typename T>
class Point_ {
public:
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    T x, y;
};
typedef Point_<float> Point2f;

template <typename T>
static inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) {
    return Point_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y));
}
template <typename T>
static inline Point_<T> &operator*=(Point_<T> &a, double b) {
    a.x = saturate_cast<T>(a.x * b);
    a.y = saturate_cast<T>(a.y * b);
    return a;
}
template <typename T>
static inline Point_<T> &operator/=(Point_<T> &a, double b) {
    a.x = saturate_cast<T>(a.x / b);
    a.y = saturate_cast<T>(a.y / b);
    return a;
}

template <typename T>
class Size_ {
public:
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T width, height;
};
typedef Size_<int> Size;

class Range {
public:
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    int start, end;
};

class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f pt_, float size_) : pt(pt_), size(size_), angle(-1), response(0), octave(0), class_id(-1) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

class Mat {
public:
    Mat() : rows(0), cols(0), data(0), step(0) {}
    // non-owning view of caller memory, like cv::Mat(rows, cols, CV_8UC1, data, step)
    Mat(int rows_, int cols_, uchar *data_, size_t step_) : rows(rows_), cols(cols_), data(data_), step(step_) {}
    void create(int rows_, int cols_) {
        rows = rows_;
        cols = cols_;
        step = (size_t)cols_;
        owner_ = std::make_shared<std::vector<uchar> >((size_t)rows_ * step + step + 2, (uchar)0);
        data = owner_->data();
    }
    int rows, cols;
    uchar *data;
    size_t step;

private:
    std::shared_ptr<std::vector<uchar> > owner_;
};

class ParallelLoopBody {
public:
    virtual ~ParallelLoopBody() {}
    virtual void operator()(const Range &range) const = 0;
};

static inline void parallel_for_(const Range &range, const ParallelLoopBody &body, double /*nstripes*/ = -1.) {
    body(range);
}

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

static inline void resize(const Mat &src, Mat &dst, Size dsize, double /*fx*/ = 0, double /*fy*/ = 0,
                          int /*interpolation*/ = INTER_LINEAR) {
    // the restated resize derives its output size from the source, int(cols * 0.5) x int(rows * 0.5): the only
    // shape the reference asks for (src/algorithm.cpp:147-150); anything else is a harness error
    if (dsize.width != (int)(src.cols * 0.5) || dsize.height != (int)(src.rows * 0.5)) std::abort();
    dst.create(dsize.height, dsize.width);
    if (klt_oracle_resize_half(src.data, src.cols, src.rows, src.step, dst.data) != 0) std::abort();
}

}  // namespace cv
#endif
