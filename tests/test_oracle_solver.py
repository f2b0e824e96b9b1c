"""Known-answer tests of the CPU oracle (the reference has no test on this path)."""
import numpy as np
import pytest

from lego_slam_b200 import synth
from oracle import klt_oracle_np as onp


def _smooth(rows, cols, seed=0):
    rng = np.random.default_rng(seed)
    n = rng.standard_normal((rows, cols))
    for _ in range(6):
        n = (n + np.roll(n, 1, 0) + np.roll(n, -1, 0) + np.roll(n, 1, 1) + np.roll(n, -1, 1)) / 5
    n -= n.min()
    return n * (255 / n.max())


# ---------------- sampler: algorithm.h:40-57 ----------------
def test_sampler_integer_coordinates_return_the_pixel(oracle):
    img = np.random.default_rng(1).integers(0, 256, size=(20, 30), dtype=np.uint8)
    for x, y in [(0, 0), (5, 7), (28, 18), (29, 19)]:
        assert oracle.get_pixel_value(img, float(x), float(y)) == float(img[y, x])


def test_sampler_clamps_only_outside(oracle):
    img = np.random.default_rng(2).integers(0, 256, size=(20, 30), dtype=np.uint8)
    assert oracle.get_pixel_value(img, -3.5, 4.0) == float(img[4, 0])
    assert oracle.get_pixel_value(img, 5.0, -1.0) == float(img[0, 5])
    assert oracle.get_pixel_value(img, 30.0, 4.0) == float(img[4, 29])   # x >= cols -> cols-1
    assert oracle.get_pixel_value(img, 5.0, 20.0) == float(img[19, 5])   # y >= rows -> rows-1


def test_sampler_sliver_uses_flat_addressing(oracle):
    """x in (cols-1, cols) is NOT clamped: tap [1] is the first byte of the next row (F7)."""
    img = np.random.default_rng(3).integers(0, 256, size=(20, 30), dtype=np.uint8)
    v = oracle.get_pixel_value(img, 29.5, 4.0)
    f = np.float32
    expect = f(0.5) * f(1) * f(img[4, 29]) + f(0.5) * f(1) * f(img[5, 0])
    assert v == float(expect)
    # y in (rows-1, rows): taps below the image read the zero padding
    v = oracle.get_pixel_value(img, 5.0, 19.5)
    assert v == float(f(0.5) * f(img[19, 5]))


def test_sampler_bilinear_value(oracle):
    img = np.array([[10, 20], [30, 40], [0, 0]], np.uint8)
    assert oracle.get_pixel_value(img, 0.25, 0.5) == pytest.approx(10 * .75 * .5 + 20 * .25 * .5 + 30 * .75 * .5 + 40 * .25 * .5)


# ---------------- 2x2 LDLT: Eigen semantics ----------------
def test_ldlt_regular(oracle):
    H = np.array([[4.0, 1.0], [1.0, 3.0]])
    b = np.array([1.0, 2.0])
    assert np.allclose(oracle.ldlt2_solve(H, b), np.linalg.solve(H, b), rtol=1e-15)


def test_ldlt_pivots_on_larger_diagonal(oracle):
    H = np.array([[1e-8, 1e-4], [1e-4, 5.0]])
    b = np.array([0.3, -0.7])
    assert np.allclose(oracle.ldlt2_solve(H, b), np.linalg.solve(H, b), rtol=1e-9)


def test_ldlt_zero_matrix_gives_zero_not_nan(oracle):
    """flat patch: H = 0 -> update 0, success stays true (F6)."""
    x = oracle.ldlt2_solve(np.zeros((2, 2)), np.array([1.0, 2.0]))
    assert np.array_equal(x, [0.0, 0.0])


def test_ldlt_rank_one(oracle):
    H = np.array([[4.0, 2.0], [2.0, 1.0]])  # d1 == 0 exactly -> second component dropped
    x = oracle.ldlt2_solve(H, np.array([2.0, 1.0]))
    assert np.all(np.isfinite(x))
    assert np.allclose(H @ x, [2.0, 1.0])


def test_ldlt_matches_numpy_restatement(oracle):
    rng = np.random.default_rng(4)
    for _ in range(200):
        a, c, d = rng.standard_normal(3) * 10.0 ** int(rng.integers(-3, 4))
        b = rng.standard_normal(2)
        got = oracle.ldlt2_solve([[a, c], [c, d]], b)
        ref = onp.ldlt2_solve(a, c, d, b[0], b[1])
        assert got[0] == ref[0] and got[1] == ref[1]


# ---------------- solver KATs: src/algorithm.cpp:37-125 ----------------
def test_pure_translation_is_recovered(oracle):
    canvas = _smooth(140, 200, 7)
    shift = (2.3, -1.4)
    yy, xx = np.mgrid[0:100, 0:160].astype(np.float64)
    a = synth._sample(canvas, xx + 20, yy + 20)
    b = synth._sample(canvas, xx + 20 - shift[0], yy + 20 - shift[1])
    a8, b8 = np.rint(a).astype(np.uint8), np.rint(b).astype(np.uint8)
    kp1 = np.array([[40, 30], [80, 50], [120, 70], [60, 60]], np.float32)
    out, succ, _ = oracle.track(a8, b8, kp1, kp1.copy(), oracle.make_params(levels=3))
    assert succ.all()
    assert np.abs(out - kp1 - np.array(shift, np.float32)).max() < 0.15


def test_flat_patch_zero_motion_success(oracle):
    img = np.full((64, 64), 128, np.uint8)
    kp = np.array([[32, 32]], np.float32)
    out, succ, st = oracle.track(img, img, kp, kp + np.float32(0.5), oracle.make_params(levels=1))
    assert succ[0] == 1 and st.n_nan == 0
    assert np.array_equal(out, kp + np.float32(0.5))  # update == 0: guess is returned unchanged


def test_point_leaving_image_fails(oracle):
    img = np.random.default_rng(8).integers(0, 256, size=(64, 64), dtype=np.uint8)
    kp1 = np.array([[10, 10], [30, 30]], np.float32)
    kp2 = np.array([[-40, 10], [30, 30]], np.float32)
    out, succ, st = oracle.track(img, img, kp1, kp2, oracle.make_params(levels=1))
    assert succ[1] == 1
    if out[0, 0] < 0:
        assert succ[0] == 0 and st.n_out_of_image >= 1


def test_has_initial_false_ignores_guess(oracle):
    L, R, kp1, kp2, _ = synth.stereo_case(94, 310, 30, seed=3, min_dist=8)
    junk = kp2 + np.float32(7.0)
    a, sa, _ = oracle.track(L, R, kp1, kp1.copy(), oracle.make_params(has_initial=False, levels=3))
    b, sb, _ = oracle.track(L, R, kp1, junk, oracle.make_params(has_initial=False, levels=3))
    assert np.array_equal(a, b) and np.array_equal(sa, sb)


def test_inverse_mode_is_the_stale_jacobian_variant(oracle):
    """Guards against 'fixing' F4: the reference's inverse mode differs from forward results."""
    L, R, kp1, kp2, _ = synth.stereo_case(94, 310, 40, seed=4, min_dist=8)
    f, _, _ = oracle.track(L, R, kp1, kp2, oracle.make_params(levels=3))
    i, _, _ = oracle.track(L, R, kp1, kp2, oracle.make_params(levels=3, inverse=True))
    assert not np.array_equal(f, i)
    p1, p2 = oracle.build_pyramid(L, 3), oracle.build_pyramid(R, 3)
    n, sn, _ = onp.track(p1, p2, kp1, kp2, levels=3, inverse=True)
    assert np.array_equal(n.view(np.uint32), i.view(np.uint32))


def test_thread_stripes_do_not_change_results(oracle):
    L, R, kp1, kp2, _ = synth.stereo_case(188, 620, 300, seed=5, min_dist=5)
    a, sa, s1 = oracle.track(L, R, kp1, kp2, threads=1)
    b, sb, s4 = oracle.track(L, R, kp1, kp2, threads=4)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(sa, sb)
    assert list(s1.gn_iters) == list(s4.gn_iters)


def test_empty_feature_list(oracle):
    img = np.zeros((32, 32), np.uint8)
    out, succ, st = oracle.track(img, img, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32),
                                 oracle.make_params(levels=2))
    assert out.shape == (0, 2) and succ.shape == (0,) and st.n_features == 0


# ---------------- frozen vectors + cross-implementation agreement ----------------
def test_oracle_reproduces_golden_vectors(oracle, golden):
    meta, vec = golden
    for name, kw in meta["solver"].items():
        kw = dict(kw)
        iters, nsucc, _src = kw.pop("gn_iters"), kw.pop("n_success"), kw.pop("source")
        p = oracle.make_params(**kw)
        out, succ, st = oracle.track(vec["left"], vec["right"], vec["kp1"], vec["kp2"], p)
        assert np.array_equal(out.view(np.uint32), vec[f"{name}_kp2"].view(np.uint32)), name
        assert np.array_equal(succ, vec[f"{name}_succ"]), name
        assert [int(v) for v in st.gn_iters][:p.levels] == iters and int(st.n_success) == nsucc, name


def test_numpy_restatement_agrees_bitwise(oracle):
    L, R, kp1, kp2, _ = synth.stereo_case(94, 310, 24, seed=6, min_dist=8)
    rng = np.random.default_rng(9)
    kp2 = (kp2 + rng.normal(0, 1.5, kp2.shape)).astype(np.float32)
    for kw in (dict(), dict(inverse=True), dict(patch_lo=-4, patch_hi=3), dict(has_initial=False)):
        p = oracle.make_params(levels=3, **kw)
        a, sa, st = oracle.track(L, R, kp1, kp2, p)
        b, sb, it = onp.track(oracle.build_pyramid(L, 3), oracle.build_pyramid(R, 3), kp1, kp2, levels=3,
                              lo=p.patch_lo, hi=p.patch_hi, inverse=bool(p.inverse),
                              has_initial=bool(p.has_initial))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), kw
        assert np.array_equal(sa, sb), kw
        assert [int(v) for v in st.gn_iters][:3] == it, kw
