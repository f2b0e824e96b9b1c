"""lego_slam_b200 -- B200-native (sm_100a) pyramid Gauss-Newton KLT tracker.

Drop-in for ONE hot path of LEGO-SLAM: legoslam::LKOpticalFlow4Layer / LKOpticalFlow1Layer
(/root/reference include/legoslam/algorithm.h:123-136, src/algorithm.cpp:11-206).  The compute path is
hand-written CUDA behind a C ABI (include/lego_klt.h); this package is the thin Python host side used
by the tests and bench.py.  It never falls back to a CPU implementation.
"""
from .api import (KERNEL_AUTO, KERNEL_EXACT, KERNEL_LANE, KERNEL_PATCH, KERNEL_WARP, Batch, Camera, Image, MultiTracker,  # noqa: F401
                  Tracker, kernel_launches, make_camera,
                  LKOpticalFlow1Layer, LKOpticalFlow4Layer, make_params, pinned_empty)
