"""Independent numpy restatement of the KLT path (TEST INFRASTRUCTURE ONLY).

Written separately from klt_oracle.cpp (vectorised over the patch, sequential sums via cumsum) so a
slip in either restatement shows up as a disagreement -- the reference has no golden vectors for this
path, so cross-implementation agreement is one of the pins we can have.

Reference text: include/legoslam/algorithm.h:40-66, src/algorithm.cpp:37-125, :128-206.
fp32 where the reference is fp32 (numpy float32 element ops are correctly rounded, no FMA), fp64
elsewhere.  Slow: use for tens of features.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def pad_plane(img: np.ndarray):
    """(flat zero-padded uint8 buffer, cols, rows, step)."""
    rows, cols = img.shape
    step = img.strides[0]
    flat = np.zeros(rows * step + step + 2, np.uint8)
    for r in range(rows):
        flat[r * step:r * step + cols] = img[r]
    return flat, cols, rows, step


def get_pixel_value(plane, x, y):
    """algorithm.h:40-57, vectorised over float32 arrays x, y."""
    flat, cols, rows, step = plane
    x = np.asarray(x, F32).copy()
    y = np.asarray(y, F32).copy()
    x[x < 0] = F32(0)
    y[y < 0] = F32(0)
    x[x >= F32(cols)] = F32(cols - 1)
    y[y >= F32(rows)] = F32(rows - 1)
    ix = x.astype(np.int64)  # truncation == int(x) for x >= 0
    iy = y.astype(np.int64)
    base = iy * step + ix
    xx = x - np.floor(x)
    yy = y - np.floor(y)
    one = F32(1)
    p0 = flat[base].astype(F32)
    p1 = flat[base + 1].astype(F32)
    p2 = flat[base + step].astype(F32)
    p3 = flat[base + step + 1].astype(F32)
    return (one - xx) * (one - yy) * p0 + xx * (one - yy) * p1 + (one - xx) * yy * p2 + xx * yy * p3


def ldlt2_solve(h00, h10, h11, b0, b1):
    """Eigen 3.3 LDLT (pivoted, pseudo-inverse of D) for 2x2; see klt_oracle.cpp."""
    perm = abs(h11) > abs(h00)
    a, d = (h11, h00) if perm else (h00, h11)
    y0, y1 = (b1, b0) if perm else (b0, b1)
    if a == 0.0:
        l, d1 = h10, d
    else:
        l = h10 / a
        d1 = d - l * (a * l)
    y1 = y1 - l * y0
    tol = 1.0 / np.finfo(np.float64).max
    y0 = y0 / a if abs(a) > tol else 0.0
    y1 = y1 / d1 if abs(d1) > tol else 0.0
    y0 = y0 - l * y1
    return (y1, y0) if perm else (y0, y1)


def _seq_sum(v):
    """left-to-right fp64 sum starting from 0.0 (same order as the reference's += loop)."""
    return float(np.cumsum(np.asarray(v, np.float64))[-1])


def track_level(plane1, plane2, kp1, kp2, lo=-3, hi=3, max_iters=10, eps=1e-2, inverse=False,
                has_initial=True):
    """src/algorithm.cpp:37-125 for every feature; returns (kp2_out, success, iters)."""
    n = kp1.shape[0]
    out = np.zeros((n, 2), F32)
    succ = np.zeros(n, np.uint8)
    iters = 0
    offs = np.arange(lo, hi + 1)
    ox = np.repeat(offs, offs.size).astype(F32)  # x outer
    oy = np.tile(offs, offs.size).astype(F32)    # y inner
    cols2, rows2 = plane2[1], plane2[2]
    for i in range(n):
        kx, ky = F32(kp1[i, 0]), F32(kp1[i, 1])
        dx = dy = 0.0
        if has_initial:
            dx = float(F32(kp2[i, 0]) - kx)
            dy = float(F32(kp2[i, 1]) - ky)
        px = kx + ox  # float32 adds
        py = ky + oy
        pxd, pyd = px.astype(np.float64), py.astype(np.float64)
        I1 = get_pixel_value(plane1, px, py)
        cost = last = 0.0
        ok = True
        H = [0.0, 0.0, 0.0]
        Jx = Jy = None
        stale = (0.0, 0.0)
        for it in range(max_iters):
            iters += 1
            cx, cy = pxd + dx, pyd + dy
            err = (I1 - get_pixel_value(plane2, cx.astype(F32), cy.astype(F32))).astype(np.float64)
            if not inverse:
                gx = get_pixel_value(plane2, (cx + 1).astype(F32), cy.astype(F32)) - \
                    get_pixel_value(plane2, (cx - 1).astype(F32), cy.astype(F32))
                gy = get_pixel_value(plane2, cx.astype(F32), (cy + 1).astype(F32)) - \
                    get_pixel_value(plane2, cx.astype(F32), (cy - 1).astype(F32))
                Jx = -1.0 * (0.5 * gx.astype(np.float64))
                Jy = -1.0 * (0.5 * gy.astype(np.float64))
            elif it == 0:
                gx = get_pixel_value(plane1, px + F32(1), py) - get_pixel_value(plane1, px - F32(1), py)
                gy = get_pixel_value(plane1, px, py + F32(1)) - get_pixel_value(plane1, px, py - F32(1))
                Jx = -1.0 * (0.5 * gx.astype(np.float64))
                Jy = -1.0 * (0.5 * gy.astype(np.float64))
                stale = (float(Jx[-1]), float(Jy[-1]))
            else:
                # the reference's J survives from the last pixel of iteration 0 (F4)
                Jx = np.full(err.shape, stale[0])
                Jy = np.full(err.shape, stale[1])
            b0 = _seq_sum(-err * Jx)
            b1 = _seq_sum(-err * Jy)
            cost = _seq_sum(err * err)
            if (not inverse) or it == 0:
                H = [_seq_sum(Jx * Jx), _seq_sum(Jy * Jx), _seq_sum(Jy * Jy)]
            u0, u1 = ldlt2_solve(H[0], H[1], H[2], b0, b1)
            if not (math.isfinite(u0) and math.isfinite(u1)):
                ok = False
                break
            if it > 0 and cost > last:
                break
            dx += u0
            dy += u1
            last = cost
            ok = True
            if math.sqrt(u0 * u0 + u1 * u1) < eps:
                break
        x2, y2 = kx + F32(dx), ky + F32(dy)
        out[i] = (x2, y2)
        inside = not (float(x2) < 0 or float(y2) < 0 or float(x2) >= cols2 or float(y2) >= rows2)
        succ[i] = 1 if (ok and inside) else 0
    return out, succ, iters


def track(pyr1, pyr2, kp1, kp2, levels=4, lo=-3, hi=3, max_iters=10, eps=1e-2, inverse=False,
          has_initial=True):
    """src/algorithm.cpp:158-205 on prebuilt pyramids (lists of uint8 level images)."""
    top = math.ldexp(1.0, -(levels - 1))
    k1 = (np.asarray(kp1, F32).astype(np.float64) * top).astype(F32)
    k2 = (np.asarray(kp2, F32).astype(np.float64) * top).astype(F32)
    planes1 = [pad_plane(p) for p in pyr1]
    planes2 = [pad_plane(p) for p in pyr2]
    per_level = [0] * levels
    succ = np.zeros(k1.shape[0], np.uint8)
    for level in range(levels - 1, -1, -1):
        hi_flag = has_initial if level == levels - 1 else True
        k2, succ, per_level[level] = track_level(planes1[level], planes2[level], k1, k2, lo, hi,
                                                 max_iters, eps, inverse, hi_flag)
        if level > 0:
            k1 = (k1.astype(np.float64) / 0.5).astype(F32)
            scaled = (k2.astype(np.float64) / 0.5).astype(F32)
            k2 = np.where(succ[:, None] != 0, scaled, k1)
    return k2, succ, per_level
