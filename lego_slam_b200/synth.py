"""Seeded synthetic KITTI-shaped inputs: stereo pairs, frame-to-frame pairs, features.

numpy only (no cv2), so the same bytes are produced here and on the GPU box.  No KITTI files exist
offline; the reference itself reads ``image_{0,1}/%06d.png`` (src/dataset.cpp:53-86).

Scene model: 3 octaves of Gaussian-blurred white noise on a canvas larger than the frame; the left
image is a crop, the right image is the same scene resampled with a row-dependent disparity
``d(y) = 5 + 30*y/H`` px (match of left x is at ``x - d``), the next frame is the scene translated by
a sub-pixel flow and scaled by 1 %; N(0,1) sensor noise; quantised to uint8.
"""
from __future__ import annotations

import numpy as np

_MARGIN_X, _MARGIN_Y = 64, 32


def _gauss_blur_fft(noise: np.ndarray, sigma: float) -> np.ndarray:
    h, w = noise.shape
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.rfftfreq(w)[None, :]
    g = np.exp(-2.0 * (np.pi * sigma) ** 2 * (fx * fx + fy * fy))
    return np.fft.irfft2(np.fft.rfft2(noise) * g, s=noise.shape)


def make_scene(rows: int, cols: int, seed: int) -> np.ndarray:
    """float64 canvas (rows+64, cols+128), values in [0,255]."""
    rng = np.random.default_rng(seed)
    h, w = rows + 2 * _MARGIN_Y, cols + 2 * _MARGIN_X
    acc = np.zeros((h, w))
    for sigma, wgt in ((1.5, 1.0), (4.0, 1.0), (12.0, 0.7)):
        acc += sigma * wgt * _gauss_blur_fft(rng.standard_normal((h, w)), sigma)
    acc -= acc.min()
    acc *= 255.0 / acc.max()
    return acc


def _sample(canvas: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    h, w = canvas.shape
    xs = np.clip(xs, 0.0, w - 1.001)
    ys = np.clip(ys, 0.0, h - 1.001)
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    fx, fy = xs - x0, ys - y0
    return ((1 - fx) * (1 - fy) * canvas[y0, x0] + fx * (1 - fy) * canvas[y0, x0 + 1]
            + (1 - fx) * fy * canvas[y0 + 1, x0] + fx * fy * canvas[y0 + 1, x0 + 1])


def _quantise(img: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    img = img + rng.standard_normal(img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def stereo_pair(rows: int = 376, cols: int = 1241, seed: int = 1):
    """(left, right, disparity_of_row) -- right(x - d(y), y) == left(x, y)."""
    canvas = make_scene(rows, cols, seed)
    rng = np.random.default_rng(seed + 7919)
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float64)
    left = _sample(canvas, xx + _MARGIN_X, yy + _MARGIN_Y)
    disp = 5.0 + 30.0 * yy / rows
    right = _sample(canvas, xx + disp + _MARGIN_X, yy + _MARGIN_Y)
    return _quantise(left, rng), _quantise(right, rng), disp[:, 0].copy()


def next_frame(rows: int, cols: int, seed: int, frame: int):
    """Left image of frame `frame` of sequence `seed`: scene translated by a sub-pixel flow
    (<= 8 px per frame) and scaled by 1 % per frame about the image centre.  Returns
    (image, (tx, ty, scale)) with  x_scene = cx + (x - cx) / scale + tx."""
    canvas = make_scene(rows, cols, seed)
    rng_flow = np.random.default_rng(seed + 104729)
    steps = rng_flow.uniform(-8.0, 8.0, size=(max(frame, 1), 2))
    tx, ty = (steps[:frame].sum(axis=0) if frame > 0 else np.zeros(2))
    # keep the crop inside the canvas margins
    tx = float(np.clip(tx, -_MARGIN_X + 8, _MARGIN_X - 8))
    ty = float(np.clip(ty, -_MARGIN_Y + 8, _MARGIN_Y - 8))
    scale = 1.01 ** frame
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float64)
    cx, cy = 0.5 * cols, 0.5 * rows
    xs = cx + (xx - cx) / scale + tx + _MARGIN_X
    ys = cy + (yy - cy) / scale + ty + _MARGIN_Y
    rng = np.random.default_rng(seed * 1000003 + frame)
    return _quantise(_sample(canvas, xs, ys), rng), (tx, ty, scale)


def _box3(a: np.ndarray) -> np.ndarray:
    p = np.pad(a, 1, mode="edge")
    return (p[:-2, :-2] + p[:-2, 1:-1] + p[:-2, 2:] + p[1:-1, :-2] + p[1:-1, 1:-1] + p[1:-1, 2:]
            + p[2:, :-2] + p[2:, 1:-1] + p[2:, 2:])


def detect_features(img: np.ndarray, n: int, min_dist: int = 5, border: int = 16,
                    seed: int = 0) -> np.ndarray:
    """Shi-Tomasi style corners (min eigenvalue of the 3x3 structure tensor) with grid-based
    minimum-distance suppression -- the role cv::GFTTDetector plays in the reference
    (src/frontend_g2o.cpp:16,279-297).  Topped up with a jittered grid if fewer than n are found.
    Returns float32 (n, 2) {x, y} with integer-valued coordinates like GFTT."""
    f = img.astype(np.float64)
    gx = np.zeros_like(f)
    gy = np.zeros_like(f)
    gx[:, 1:-1] = 0.5 * (f[:, 2:] - f[:, :-2])
    gy[1:-1, :] = 0.5 * (f[2:, :] - f[:-2, :])
    a, b, c = _box3(gx * gx), _box3(gx * gy), _box3(gy * gy)
    score = 0.5 * (a + c) - np.sqrt(0.25 * (a - c) ** 2 + b * b)
    rows, cols = img.shape
    score[:border, :] = score[-border:, :] = 0
    score[:, :border] = score[:, -border:] = 0
    # local maxima over 3x3
    p = np.pad(score, 1, mode="constant")
    nb = np.stack([p[1 + dy:1 + dy + rows, 1 + dx:1 + dx + cols]
                   for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dy, dx) != (0, 0)])
    ismax = (score > 0) & (score >= nb.max(axis=0))
    ys, xs = np.nonzero(ismax)
    order = np.argsort(-score[ys, xs], kind="stable")
    ys, xs = ys[order], xs[order]
    cell = max(int(min_dist), 1)
    gh, gw = rows // cell + 3, cols // cell + 3
    occ = np.zeros((gh, gw), dtype=bool)
    picked = []
    for y, x in zip(ys.tolist(), xs.tolist()):
        cy, cx = y // cell + 1, x // cell + 1
        if occ[cy - 1:cy + 2, cx - 1:cx + 2].any():
            continue
        occ[cy, cx] = True
        picked.append((x, y))
        if len(picked) >= n:
            break
    if len(picked) < n:  # jittered-grid top-up
        rng = np.random.default_rng(seed + 15485863)
        need = n - len(picked)
        gx_ = rng.integers(border, cols - border, size=need)
        gy_ = rng.integers(border, rows - border, size=need)
        picked.extend(zip(gx_.tolist(), gy_.tolist()))
    return np.asarray(picked[:n], dtype=np.float32).reshape(n, 2)


def stereo_case(rows: int = 376, cols: int = 1241, n: int = 150, seed: int = 1, min_dist=None,
                guess: str = "same"):
    """One left->right KLT problem: (left, right, kp1, kp2_init, kp2_truth).
    guess='same' -> kp2 = kp1 (src/frontend_g2o.cpp:508, unmapped features);
    guess='noisy' -> truth + N(0, 2 px) (the projected-map-point branch, :504-505)."""
    left, right, disp = stereo_pair(rows, cols, seed)
    if min_dist is None:
        min_dist = 20 if n <= 300 else 5
    kp1 = detect_features(left, n, min_dist=min_dist, seed=seed)
    truth = kp1.copy()
    truth[:, 0] -= disp[kp1[:, 1].astype(np.int64)].astype(np.float32)
    if guess == "same":
        kp2 = kp1.copy()
    else:
        rng = np.random.default_rng(seed + 32452843)
        kp2 = (truth + rng.normal(0, 2.0, size=truth.shape)).astype(np.float32)
    return left, right, kp1, kp2, truth


def temporal_case(rows: int = 376, cols: int = 1241, n: int = 2000, seed: int = 2, frame: int = 1,
                  guess: str = "same"):
    """Frame-to-frame problem (Frontend::TrackLastFrame): last-left -> current-left."""
    prev, (tx0, ty0, s0) = next_frame(rows, cols, seed, frame - 1)
    cur, (tx1, ty1, s1) = next_frame(rows, cols, seed, frame)
    kp1 = detect_features(prev, n, min_dist=5 if n > 300 else 20, seed=seed + frame)
    cx, cy = 0.5 * cols, 0.5 * rows
    # scene coordinate of each feature, then its pixel in the current frame
    xs = cx + (kp1[:, 0] - cx) / s0 + tx0
    ys = cy + (kp1[:, 1] - cy) / s0 + ty0
    truth = np.stack([cx + (xs - tx1 - cx) * s1, cy + (ys - ty1 - cy) * s1], axis=1).astype(np.float32)
    if guess == "same":
        kp2 = kp1.copy()
    else:
        rng = np.random.default_rng(seed + 49979687 + frame)
        kp2 = (truth + rng.normal(0, 2.0, size=truth.shape)).astype(np.float32)
    return prev, cur, kp1, kp2, truth
