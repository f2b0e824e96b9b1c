"""The oracle restatement against oracle/_ref -- the reference's OWN translation unit (src/algorithm.cpp +
include/legoslam/algorithm.h compiled unmodified on stand-in headers, oracle/build_ref.py).  Bitwise: positions and
flags.  This is the pin of the solver part of the oracle; the pyramid part is pinned to cv2."""
import numpy as np
import pytest

from lego_slam_b200 import synth
from oracle import binding as ob
from oracle import ref_binding as rb

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built and /root/reference absent")


def _ugly(cols, rows):
    """border / sliver / outside / binade-boundary / sub-pixel source points"""
    return np.array([[3.0, 3.0], [cols - 2.5, rows - 2.5], [cols - 0.5, 40.0], [100.0, rows - 0.25], [-2.0, 50.0],
                     [127.999, 63.999], [255.5, 31.75], [511.9996, 100.0], [0.0, 0.0], [cols - 1.0, rows - 1.0],
                     [cols + 3.0, rows + 3.0], [1.25, rows - 1.5], [cols - 1.25, 2.5], [64.0, 0.5]], np.float32)


def _case(rows, cols, n, seed, guess="same", subpixel=False):
    L, R, kp1, kp2, _ = synth.stereo_case(rows, cols, n, seed=seed, min_dist=5 if n > 300 else 10, guess=guess)
    if subpixel:
        rng = np.random.default_rng(seed + 5)
        kp1 = (kp1 + rng.uniform(-0.5, 0.5, kp1.shape)).astype(np.float32)
        kp2 = (kp2 + rng.uniform(-0.5, 0.5, kp2.shape)).astype(np.float32)
    ex = _ugly(cols, rows)
    kp1 = np.concatenate([kp1, ex]).astype(np.float32)
    kp2 = np.concatenate([kp2, ex + np.float32(0.3)]).astype(np.float32)
    return L, R, kp1, kp2


def _same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("has_initial", [True, False])
@pytest.mark.parametrize("shape,n,seed,guess,subpixel", [
    ((376, 1241), 150, 1, "same", False),      # C1
    ((376, 1241), 2000, 2, "same", False),     # C2 / C3 pair shape
    ((376, 1241), 600, 1003, "noisy", True),   # projected-guess branch, tracked sub-pixel source points
    ((188, 620), 150, 1, "same", False),       # what the reference pipeline really feeds (F11)
    ((97, 131), 60, 5, "noisy", True),         # small odd shape: many border cases
])
def test_oracle_equals_reference_tu_4layer(shape, n, seed, guess, subpixel, inverse, has_initial):
    L, R, kp1, kp2 = _case(shape[0], shape[1], n, seed, guess, subpixel)
    ref_kp, ref_ok = rb.track(L, R, kp1, kp2, inverse=inverse, has_initial=has_initial)
    o_kp, o_ok, _ = ob.track(L, R, kp1, kp2, ob.make_params(inverse=inverse, has_initial=has_initial))
    assert np.array_equal(o_ok, ref_ok)
    assert _same(o_kp, ref_kp)


@pytest.mark.parametrize("inverse", [False, True])
def test_oracle_equals_reference_tu_1layer(inverse):
    L, R, kp1, kp2 = _case(188, 620, 200, 7, "noisy", True)
    ref_kp, ref_ok = rb.track(L, R, kp1, kp2, inverse=inverse, layers=1)
    o_kp, o_ok, _ = ob.track(L, R, kp1, kp2, ob.make_params(levels=1, inverse=inverse))
    assert np.array_equal(o_ok, ref_ok) and _same(o_kp, ref_kp)


def test_oracle_equals_reference_tu_padded_step():
    """cv::Mat with step > cols (a ROI / aligned rows): flat addressing reads the padding bytes (F7)."""
    L, R, kp1, kp2 = _case(120, 200, 80, 9, "noisy", True)
    rng = np.random.default_rng(0)
    Lp = rng.integers(0, 256, size=(120, 224), dtype=np.uint8)
    Rp = rng.integers(0, 256, size=(120, 224), dtype=np.uint8)
    Lp[:, :200], Rp[:, :200] = L, R
    ref_kp, ref_ok = rb.track(Lp[:, :200], Rp[:, :200], kp1, kp2)
    o_kp, o_ok, _ = ob.track(Lp[:, :200], Rp[:, :200], kp1, kp2)
    assert np.array_equal(o_ok, ref_ok) and _same(o_kp, ref_kp)


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("half_patch,pyramids,shape,n,seed", [
    (3, 5, (1080, 1920), 1200, 3),     # C4 shape: 5 levels (literal-substituted reference text)
    (5, 4, (376, 1241), 400, 4),       # C5: 11x11 patch
    (5, 3, (188, 620), 158, 1),        # the 11x11 golden-vector configuration
])
def test_oracle_equals_parametrised_reference_tu(half_patch, pyramids, shape, n, seed, inverse):
    L, R, kp1, kp2 = _case(shape[0], shape[1], n, seed, "noisy", True)
    ref_kp, ref_ok = rb.track(L, R, kp1, kp2, inverse=inverse, half_patch=half_patch, pyramids=pyramids)
    p = ob.make_params(levels=pyramids, patch_lo=-half_patch, patch_hi=half_patch, inverse=inverse)
    o_kp, o_ok, _ = ob.track(L, R, kp1, kp2, p)
    assert np.array_equal(o_ok, ref_ok) and _same(o_kp, ref_kp)


def test_flat_patch_and_nan_paths():
    """flat image: H = 0 -> Eigen's pivoted LDLT gives update 0, success stays true (F6); identical in both."""
    L = np.full((64, 96), 77, np.uint8)
    R = np.full((64, 96), 200, np.uint8)
    kp1 = np.array([[20.0, 20.0], [50.5, 30.25], [94.0, 62.0]], np.float32)
    ref_kp, ref_ok = rb.track(L, R, kp1, kp1 + np.float32(1.5))
    o_kp, o_ok, _ = ob.track(L, R, kp1, kp1 + np.float32(1.5))
    assert np.array_equal(o_ok, ref_ok) and _same(o_kp, ref_kp)
    assert ref_ok.all()


def test_random_textures_many_seeds():
    """white-noise images (worst case for GN: many cost-increase exits, pivots on either diagonal entry)."""
    for seed in range(6):
        rng = np.random.default_rng(seed)
        L = rng.integers(0, 256, size=(90, 140), dtype=np.uint8)
        R = np.roll(L, 2, axis=1)
        kp1 = rng.uniform(-2, 142, size=(300, 2)).astype(np.float32)
        kp1[:, 1] = rng.uniform(-2, 92, size=300).astype(np.float32)
        kp2 = (kp1 + rng.normal(0, 1.5, kp1.shape)).astype(np.float32)
        for inverse in (False, True):
            ref_kp, ref_ok = rb.track(L, R, kp1, kp2, inverse=inverse)
            o_kp, o_ok, _ = ob.track(L, R, kp1, kp2, ob.make_params(inverse=inverse))
            assert np.array_equal(o_ok, ref_ok) and _same(o_kp, ref_kp), (seed, inverse)


def test_track_pairs_equals_single_calls():
    imgs1, imgs2, k1, k2 = [], [], [], []
    for s in range(3):
        L, R, kp1, kp2 = _case(97, 131, 40, 20 + s, "noisy", True)
        imgs1.append(L); imgs2.append(R); k1.append(kp1); k2.append(kp2)
    out, ok = rb.track_pairs(np.stack(imgs1), np.stack(imgs2), np.stack(k1), np.stack(k2), threads=3)
    for s in range(3):
        a, b = rb.track(imgs1[s], imgs2[s], k1[s], k2[s])
        assert _same(out[s], a) and np.array_equal(ok[s], b)


def _fuzz_image(rng, rows, cols, kind):
    if kind == 0:      # white noise
        return rng.integers(0, 256, size=(rows, cols), dtype=np.uint8)
    if kind == 1:      # smooth texture
        yy, xx = np.mgrid[0:rows, 0:cols]
        f = rng.uniform(0.05, 0.6, 4)
        img = 127 + 60 * np.sin(f[0] * xx + f[1] * yy) + 60 * np.cos(f[2] * xx - f[3] * yy)
        return np.clip(img + rng.normal(0, 3, img.shape), 0, 255).astype(np.uint8)
    if kind == 2:      # constant (singular normal equations everywhere)
        return np.full((rows, cols), int(rng.integers(0, 256)), np.uint8)
    img = np.zeros((rows, cols), np.uint8)     # step edges: rank-1 normal equations
    img[:, cols // 2:] = 200
    return img


@pytest.mark.parametrize("seed", range(40))
def test_oracle_equals_reference_tu_on_random_problems(seed):
    """Seeded random problems: image size, texture (noise / smooth / constant / step edge), keypoints inside, on the
    border, outside, far and absurdly far outside, random guesses, either mode, with and without the initial guess,
    1 or 4 layers.  Positions compared by bits (NaN payloads included), flags exactly."""
    rng = np.random.default_rng(20_000 + seed)
    rows, cols = int(rng.integers(40, 160)), int(rng.integers(40, 220))
    kind = int(rng.integers(0, 4))
    img1 = _fuzz_image(rng, rows, cols, kind)
    img2 = np.roll(img1, (int(rng.integers(-2, 3)), int(rng.integers(-3, 4))), axis=(0, 1)) if rng.random() < 0.7 \
        else _fuzz_image(rng, rows, cols, int(rng.integers(0, 4)))
    n = int(rng.integers(1, 50))
    kp1 = np.stack([rng.uniform(-4, cols + 4, n), rng.uniform(-4, rows + 4, n)], axis=1).astype(np.float32)
    if rng.random() < 0.5:
        kp1 = np.round(kp1)                      # freshly detected corners are integer
    kp2 = (kp1 + rng.normal(0, 2.5, kp1.shape)).astype(np.float32)
    # (finite only: a NaN coordinate indexes the image with int(NaN) in the reference -- undefined behaviour,
    # algorithm.h:42-48 clamps only x < 0 and x >= cols -- and +-inf turns into NaN through k2 - k1)
    special = np.float32([1e7, -1e7, 3e38, -3e38, 65536.5, 0.0, -0.0, 1e-40])
    for _ in range(int(rng.integers(0, 5))):
        (kp1 if rng.random() < 0.5 else kp2)[int(rng.integers(0, n)), int(rng.integers(0, 2))] = special[int(rng.integers(0, len(special)))]
    inverse, has_initial = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    layers = 1 if rng.random() < 0.25 else 4
    with np.errstate(all="ignore"):
        ref_kp, ref_ok = rb.track(img1, img2, kp1, kp2, inverse=inverse, has_initial=has_initial, layers=layers)
        o_kp, o_ok, _ = ob.track(img1, img2, kp1, kp2, ob.make_params(levels=layers, inverse=inverse, has_initial=has_initial))
    assert np.array_equal(o_ok, ref_ok)
    assert _same(o_kp, ref_kp)
