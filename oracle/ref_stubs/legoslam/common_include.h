// Stand-in for the reference's include/legoslam/common_include.h (TEST INFRASTRUCTURE ONLY -- oracle/_ref build).
//
// The real header pulls Eigen, Sophus, OpenCV and glog, none of which exist in this image.  This one is found first
// on the include path, so that the reference's include/legoslam/algorithm.h and src/algorithm.cpp compile UNMODIFIED
// where they lie.  It provides the names those two files use:
//   - Eigen::Matrix2d / Vector2d with the operations of src/algorithm.cpp:55-57,59,61,68-87,92-93,113, implemented in
//     eigen_stub.hpp (incl. a restatement of Eigen 3.3's pivoted LDLT -- third party, see that file);
//   - the types legoslam::triangulation (algorithm.h:11-34) mentions, DECLARED ONLY: that inline function is never
//     called by the hot path, so it only has to parse; no code is generated for it.
#ifndef LEGO_REF_STUB_COMMON_INCLUDE_H
#define LEGO_REF_STUB_COMMON_INCLUDE_H

#include <cmath>
#include <iostream>
#include <vector>

#include "eigen_stub.hpp"
#include <opencv2/opencv.hpp>

typedef Eigen::ref_stub::Loose MatXX;
typedef Eigen::ref_stub::Loose VecX;
typedef Eigen::ref_stub::Loose Mat34;
typedef Eigen::ref_stub::LooseVec3 Vec3;
typedef Eigen::Vector2d Vec2;
typedef std::vector<Vec3> VecVec3;
typedef Eigen::ref_stub::LooseSE3 SE3;

using cv::Mat;

#endif
