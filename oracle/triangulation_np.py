"""CPU restatement of legoslam::triangulation (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Follows include/legoslam/algorithm.h:11-34 line by line:
    :15-22  A (2n x 4): A[2i] = points[i][0] * m.row(2) - m.row(0);  A[2i+1] = points[i][1] * m.row(2) - m.row(1)
    :23     svd = A.bdcSvd(ComputeThinU | ComputeThinV)
    :24     pt_world = (V.col(3) / V(3,3)).head<3>()
    :26-29  NaN / Inf -> false
    :31-34  return S[3] / S[2] < singRatioThr

Third-party arithmetic: Eigen's bdcSvd (Eigen3, un-pinned, not under /root/reference; JacobiSVD below 16 columns).
It is restated by its published contract -- A = U S V^T with S descending -- through LAPACK (numpy.linalg.svd):
singular values and the null-space direction are unique, and V.col(3) / V(3,3) does not depend on the sign of
the column, so the result is defined up to rounding.  Pin: the reference's own known-answer test
test/legoslam_test_triangulation.cpp:5-23 (tests/test_triangulation.py restates it).

Also restated: Camera::pixel2camera (src/camera.cpp:21-25) and the quaternion / SE3 construction the
reference's test uses (Eigen::Quaterniond(w, x, y, z), Sophus::SE3d(q, t)).
"""
from __future__ import annotations

import numpy as np


def triangulation(poses34: np.ndarray, points: np.ndarray, sing_ratio_thr: float = 1e-3):
    """poses34: (n_views, 3, 4); points: (n_views, >=2).  Returns (pt_world (3,), ok)."""
    poses34 = np.asarray(poses34, np.float64).reshape(-1, 3, 4)
    points = np.asarray(points, np.float64)
    n = poses34.shape[0]
    A = np.zeros((2 * n, 4))
    for i in range(n):
        m = poses34[i]
        A[2 * i] = points[i][0] * m[2] - m[0]
        A[2 * i + 1] = points[i][1] * m[2] - m[1]
    _, S, Vt = np.linalg.svd(A, full_matrices=False)
    V = Vt.T
    with np.errstate(all="ignore"):
        pt = (V[:, 3] / V[3, 3])[:3]
        if not np.all(np.isfinite(pt)):
            return pt, False
        return pt, bool(S[3] / S[2] < sing_ratio_thr)


def triangulation_batch(poses34, points_xy, sing_ratio_thr: float = 1e-3):
    """points_xy: (n, n_views, 2).  Returns (pt_world (n,3), ok (n,) uint8, ratio (n,))."""
    poses34 = np.asarray(poses34, np.float64).reshape(-1, 3, 4)
    pts = np.asarray(points_xy, np.float64)
    n, nv = pts.shape[0], poses34.shape[0]
    A = np.empty((n, 2 * nv, 4))
    for i in range(nv):
        m = poses34[i]
        A[:, 2 * i] = pts[:, i, 0:1] * m[2] - m[0]
        A[:, 2 * i + 1] = pts[:, i, 1:2] * m[2] - m[1]
    _, S, Vt = np.linalg.svd(A, full_matrices=False)
    with np.errstate(all="ignore"):
        pt = Vt[:, 3, :3] / Vt[:, 3, 3:4]
        ratio = S[:, 3] / S[:, 2]
        ok = np.all(np.isfinite(pt), axis=1) & (ratio < sing_ratio_thr)
    return pt, ok.astype(np.uint8), ratio


def pixel2camera(kp_xy, fx, fy, cx, cy):
    """Camera::pixel2camera, depth 1 (src/camera.cpp:21-25): ((u - cx) * 1 / fx, (v - cy) * 1 / fy)."""
    kp = np.asarray(kp_xy, np.float32).astype(np.float64)
    return np.stack([(kp[..., 0] - cx) * 1.0 / fx, (kp[..., 1] - cy) * 1.0 / fy], axis=-1)


def quat_to_rot(w, x, y, z):
    """Rotation matrix of the (normalised) quaternion -- Eigen::Quaterniond(w, x, y, z).toRotationMatrix()."""
    q = np.array([w, x, y, z], np.float64)
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def se3_matrix3x4(rot, t):
    return np.hstack([np.asarray(rot, np.float64), np.asarray(t, np.float64).reshape(3, 1)])
