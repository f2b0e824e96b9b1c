"""CPU restatement of the feature detector of the reference's frontend (TEST INFRASTRUCTURE ONLY).

Reference: Frontend::DetectFeatures, /root/reference src/frontend_g2o.cpp:279-297 --
    gftt_ = cv::GFTTDetector::create(num_features, 0.01, 20)              (:16)
    mask  = 255 everywhere, 0 in the rectangle pt +- (10, 10) around every existing left feature   (:280-284)
    gftt_->detect(left_img, keypoints, mask)                                                          (:287)
THIRD PARTY, RESTATED: OpenCV imgproc featureselect.cpp (goodFeaturesToTrack) and corner.cpp (cornerMinEigenVal):
  1. Dx, Dy = Sobel 3x3 of the 8-bit image in fp32, scaled by 1 / (4 * blockSize * 255) = 1/3060, BORDER_REFLECT_101
  2. cov = (Dx*Dx, Dx*Dy, Dy*Dy); 3x3 box SUMS of cov (unnormalised), BORDER_REFLECT_101
  3. eig = (a + c) - sqrt((a - c)^2 + b^2)  with  a = cov0 * 0.5, b = cov1, c = cov2 * 0.5
  4. maxVal = max(eig) over the mask; eig = eig > maxVal * qualityLevel ? eig : 0; tmp = 3x3 dilation of eig
  5. candidates: 1 <= y < rows-1, 1 <= x < cols-1, eig != 0, eig == tmp, mask != 0
  6. sorted by value, descending (ties: higher address first)
  7. greedy: a candidate is kept unless an already kept corner lies closer than minDistance; stop at maxCorners
PINNING: tests/test_gftt.py compares with Python cv2 (4.13 in this image).  OpenCV's fp32 box filter sums in an order
(running sums, SIMD / IPP) that its published behaviour does not fix, so the pin is tolerance-aware:
  * eigenvalue map: |ours - cv2| <= EIG_TOL_REL * (a + c)   (the cancellation in step 3 is relative to a + c)
  * corner list: equal to cv2's, except for corners whose decision is AMBIGUOUS at that tolerance -- a candidate /
    non-candidate or a pair of competing corners (closer than minDistance) whose scores differ by less than the
    tolerance; those are reported, and must be few.
"""
from __future__ import annotations

import numpy as np

EIG_TOL_REL = 8e-6   # relative to the trace a + c of the 3x3 structure tensor


def reflect101(idx, n):
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * (n - 1) - idx, idx)


def _shift(a, dy, dx):
    """a[y + dy, x + dx] with BORDER_REFLECT_101."""
    rows, cols = a.shape
    ys = reflect101(np.arange(rows) + dy, rows)
    xs = reflect101(np.arange(cols) + dx, cols)
    return a[np.ix_(ys, xs)]


def sobel_scaled(img: np.ndarray):
    """(Dx, Dy) fp32: row pass (p[x+1] - p[x-1]) * s resp. p[x-1] + 2 p[x] + p[x+1], column pass the other kernel."""
    f32 = np.float32
    p = img.astype(f32)
    s = f32(1.0 / 3060.0)
    rx = (_shift(p, 0, 1) - _shift(p, 0, -1)) * s            # derivative along x, scaled
    dx = (_shift(rx, -1, 0) + _shift(rx, 0, 0) * f32(2)) + _shift(rx, 1, 0)
    sx = (_shift(p, 0, -1) + p * f32(2)) + _shift(p, 0, 1)   # smoothing along x
    dy = (_shift(sx, 1, 0) - _shift(sx, -1, 0)) * s
    return dx.astype(f32), dy.astype(f32)


def structure_sums(img: np.ndarray):
    """3x3 box sums of (Dx^2, Dx*Dy, Dy^2), fp32, rows summed first."""
    dx, dy = sobel_scaled(img)
    out = []
    for c in (dx * dx, dx * dy, dy * dy):
        r = (_shift(c, 0, -1) + c) + _shift(c, 0, 1)
        out.append(((_shift(r, -1, 0) + r) + _shift(r, 1, 0)).astype(np.float32))
    return out


def corner_min_eigen_val(img: np.ndarray):
    """(eig, trace) fp32: cv::cornerMinEigenVal(img, eig, 3, 3) and a + c (the scale of its rounding error)."""
    f32 = np.float32
    sxx, sxy, syy = structure_sums(img)
    a, b, c = sxx * f32(0.5), sxy, syy * f32(0.5)
    eig = (a + c) - np.sqrt((a - c) * (a - c) + b * b, dtype=f32)
    return eig.astype(f32), (a + c).astype(f32)


def exclusion_mask(rows, cols, exclude_xy, half=10.0):
    """cv::rectangle(mask, pt - (h, h), pt + (h, h), 0, FILLED): Point2f -> Point rounds half to even; both corners
    are inside the filled rectangle."""
    mask = np.full((rows, cols), 255, np.uint8)
    for x, y in np.asarray(exclude_xy, np.float32).reshape(-1, 2):
        x1, y1 = int(np.rint(np.float32(x - half))), int(np.rint(np.float32(y - half)))
        x2, y2 = int(np.rint(np.float32(x + half))), int(np.rint(np.float32(y + half)))
        mask[max(y1, 0):max(y2 + 1, 0), max(x1, 0):max(x2 + 1, 0)] = 0
    return mask


def candidates(eig: np.ndarray, mask, quality_level: float):
    """Steps 4-6: (ys, xs, values) sorted by value descending, ties by flat index descending."""
    rows, cols = eig.shape
    m = np.ones_like(eig, bool) if mask is None else mask != 0
    if not m.any():
        return np.zeros(0, int), np.zeros(0, int), np.zeros(0, np.float32)
    max_val = np.float64(eig[m].max())
    thr = np.float32(max_val * quality_level)          # cv::threshold on a CV_32F image compares with (float)thresh
    e = np.where(eig > thr, eig, np.float32(0))
    pad = np.pad(e, 1, mode="constant", constant_values=-np.inf)
    dil = np.max(np.stack([pad[1 + dy:1 + dy + rows, 1 + dx:1 + dx + cols] for dy in (-1, 0, 1) for dx in (-1, 0, 1)]), axis=0)
    ok = (e != 0) & (e == dil) & m
    ok[0, :] = ok[-1, :] = False
    ok[:, 0] = ok[:, -1] = False
    ys, xs = np.nonzero(ok)
    vals = e[ys, xs]
    order = np.lexsort((-(ys * cols + xs), -vals.astype(np.float64)))
    return ys[order], xs[order], vals[order]


def greedy_min_distance(ys, xs, max_corners: int, min_distance: float):
    """Step 7.  Returns the indices (into the sorted candidate arrays) of the kept corners, in order."""
    kept = []
    if min_distance < 1:
        n = len(ys) if max_corners <= 0 else min(len(ys), max_corners)
        return list(range(n))
    cell = max(int(np.rint(min_distance)), 1)
    grid = {}
    md2 = min_distance * min_distance
    for i, (y, x) in enumerate(zip(ys.tolist(), xs.tolist())):
        cy, cx = y // cell, x // cell
        good = True
        for gy in (cy - 1, cy, cy + 1):
            for gx in (cx - 1, cx, cx + 1):
                for (py, px) in grid.get((gy, gx), ()):
                    if (px - x) ** 2 + (py - y) ** 2 < md2:
                        good = False
                        break
                if not good:
                    break
            if not good:
                break
        if good:
            grid.setdefault((cy, cx), []).append((y, x))
            kept.append(i)
            if max_corners > 0 and len(kept) == max_corners:
                break
    return kept


def good_features_to_track(img: np.ndarray, max_corners: int, quality_level: float = 0.01, min_distance: float = 20.0,
                           mask=None):
    """cv::goodFeaturesToTrack(img, corners, max_corners, quality_level, min_distance, mask, 3, false).
    Returns (corners float32 (n, 2) {x, y}, scores float32 (n,))."""
    eig, _ = corner_min_eigen_val(img)
    ys, xs, vals = candidates(eig, mask, quality_level)
    keep = greedy_min_distance(ys, xs, max_corners, min_distance)
    pts = np.stack([xs[keep], ys[keep]], axis=1).astype(np.float32) if keep else np.zeros((0, 2), np.float32)
    return pts, vals[keep].astype(np.float32) if keep else np.zeros(0, np.float32)


def compare_corner_lists(ours, theirs, eig, trace, min_distance: float, tol_rel: float = EIG_TOL_REL):
    """Tolerance-aware comparison of two corner lists on one image (see the module docstring).  A corner present in only
    one list is EXPLAINED if its own score is within tolerance of the quality threshold region of doubt -- taken as: some
    corner of the union of both lists lies closer than min_distance and their scores differ by at most the tolerance --
    or if it is the last corner of a truncated list.  Returns a report dict."""
    a = {(int(x), int(y)) for x, y in np.asarray(ours).reshape(-1, 2)}
    b = {(int(x), int(y)) for x, y in np.asarray(theirs).reshape(-1, 2)}
    union = list(a | b)
    only = list(a ^ b)
    tol = lambda p: tol_rel * float(trace[p[1], p[0]]) * 4.0   # noqa: E731  (a few operations deep)
    unexplained = []
    for p in only:
        sp = float(eig[p[1], p[0]])
        ok = False
        for q in union:
            if q == p:
                continue
            if (q[0] - p[0]) ** 2 + (q[1] - p[1]) ** 2 < min_distance * min_distance:
                if abs(float(eig[q[1], q[0]]) - sp) <= max(tol(p), tol(q)):
                    ok = True
                    break
        if not ok:
            unexplained.append(p)
    return {"n_ours": len(a), "n_theirs": len(b), "n_common": len(a & b), "n_only_one_list": len(only),
            "unexplained": unexplained}
