"""CPU restatement of the frame halving in Dataset::NextFrame (TEST INFRASTRUCTURE ONLY).

/root/reference src/dataset.cpp:75-77:
    cv::resize(image_left, image_left_resized, cv::Size(), 0.5, 0.5, cv::INTER_NEAREST);
OpenCV is third party (not under /root/reference; `find_package(OpenCV 3.2)`, un-pinned).  Its published
behaviour for this call, restated: dsize = (cvRound(cols * 0.5), cvRound(rows * 0.5)) with cvRound = round half
to even; dst(y, x) = src(min(floor(y * 2), rows - 1), min(floor(x * 2), cols - 1)).  Pinned bit-exact against
Python cv2 (4.13, the only OpenCV in this image) by tests/test_oracle_pyramid.py.
"""
import numpy as np


def half_size(v: int) -> int:
    """cvRound(v * 0.5): round half to even."""
    return int(np.rint(v * 0.5))


def downscale_half_nearest(img: np.ndarray) -> np.ndarray:
    rows, cols = img.shape
    ys = np.minimum(np.arange(half_size(rows)) * 2, rows - 1)
    xs = np.minimum(np.arange(half_size(cols)) * 2, cols - 1)
    return np.ascontiguousarray(img[np.ix_(ys, xs)])
