"""GPU parity tests proper: every call goes through the C ABI (liblego_klt.so) and is compared with
the CPU oracle on the same seeded inputs.  Tolerances (BASELINE.json north_star): success flags
bit-identical (borderline features reported), positions within 1e-3 px; the EXACT kernel and the
pyramid are held to bit-exactness."""
import numpy as np
import pytest

import lego_slam_b200 as klt
from lego_slam_b200 import synth
from parity_util import assert_parity

pytestmark = pytest.mark.gpu

FAST_KERNELS = [klt.KERNEL_WARP, klt.KERNEL_LANE, klt.KERNEL_PATCH]


def _supported(kernel, kw):
    """The LANE kernel is compiled for the reference's patch (7x7, -3..3), the 8x8 patch of the metric text (-4..3)
    and the 11x11 stress patch (-5..5), forward mode."""
    if kernel != klt.KERNEL_LANE:
        return True
    patch = (kw.get("patch_lo", -3), kw.get("patch_hi", 3))
    return patch == (-3, 3) if kw.get("inverse", False) else patch in ((-3, 3), (-4, 3), (-5, 5))


def _iters(st, levels):
    return [int(v) for v in st.gn_iters][:levels]


# ------------------------------------------------------------------ pyramid (K1): bit exact
@pytest.mark.parametrize("rows,cols,levels", [(376, 1241, 4), (188, 620, 4), (1080, 1920, 5), (217, 333, 3),
                                              (95, 157, 4), (64, 64, 5), (9, 11, 2), (376, 1241, 1),
                                              # levels 2 and 3 both exact halvings (the fused streaming kernel): level-3
                                              # width even / odd, a single level-3 row, a padded step below
                                              (720, 1280, 4), (1080, 1920, 4), (8, 24, 4), (16, 1241, 4), (377, 1243, 4)])
def test_pyramid_bit_exact(tracker, oracle, rows, cols, levels):
    img = np.random.default_rng(rows * 31 + cols).integers(0, 256, size=(rows, cols), dtype=np.uint8)
    got = tracker.build_pyramid(img, levels)
    ref = oracle.build_pyramid(img, levels)
    for l in range(1, levels):
        assert got[l].shape == ref[l].shape
        assert np.array_equal(got[l], ref[l]), f"level {l}: {(got[l] != ref[l]).sum()} pixels differ"


def test_pyramid_with_row_step(tracker, oracle):
    big = np.random.default_rng(3).integers(0, 256, size=(120, 400), dtype=np.uint8)
    view = big[:, :333]
    got, ref = tracker.build_pyramid(view, 3), oracle.build_pyramid(view, 3)
    for l in range(1, 3):
        assert np.array_equal(got[l], ref[l])


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,levels,pad", [(376, 1241, 4, 0), (270, 481, 5, 0), (120, 333, 3, 67), (97, 203, 1, 0),
                                                  (1080, 1920, 5, 0), (720, 1280, 4, 0), (376, 1241, 4, 39), (8, 24, 4, 0),
                                                  (17, 1243, 4, 5)])
def test_row_aprons_of_every_level(tracker, oracle, rows, cols, levels, pad):
    """The device layout the solver's border path relies on (LevelView): 32 bytes left of column 0 replicate it,
    column `cols` holds the reference's flat-address neighbour data[r*step + cols] (algorithm.h:48,53: first pixel
    of the next row, 0 after the last row; inside the row padding when step > cols), the rest replicates the last
    pixel.  Written by the fused pyramid kernel (bands of neighbouring CTAs hand the wrap byte over)."""
    big = np.random.default_rng(rows + cols).integers(0, 256, size=(rows, cols + pad), dtype=np.uint8)
    img = big[:, :cols]
    tracker.build_pyramid(img, levels)
    ref = oracle.build_pyramid(np.ascontiguousarray(img), levels)
    for l in range(levels):
        raw, left = tracker.debug_read_level(l, ref[l].shape[0])
        r, c = ref[l].shape
        pitch = raw.shape[1]
        assert np.array_equal(raw[:, left:left + c], ref[l]), f"level {l} pixels"
        assert np.array_equal(raw[:, :left], np.repeat(ref[l][:, :1], left, axis=1)), f"level {l} left apron"
        if l == 0 and pad:
            wrap = big[:, cols].copy()
            wrap[-1] = 0       # the ABI defines everything after the last pixel of the image as 0 (oracle: zero padding)
        else:
            wrap = np.concatenate([ref[l][1:, 0], [0]]).astype(np.uint8)
        assert np.array_equal(raw[:, left + c], wrap), f"level {l} wrap byte"
        n_right = pitch - left - c - 1 - left          # the last `left` bytes of a row are the next row's left apron
        assert n_right >= 15
        assert np.array_equal(raw[:, left + c + 1:left + c + 1 + n_right],
                              np.repeat(ref[l][:, -1:], n_right, axis=1)), f"level {l} right apron"


# ------------------------------------------------------------------ golden vectors through the C ABI
def test_exact_kernel_reproduces_golden_vectors_bitwise(tracker, golden):
    meta, vec = golden
    for name, kw in meta["solver"].items():
        kw = dict(kw)
        iters, nsucc, _src = kw.pop("gn_iters"), kw.pop("n_success"), kw.pop("source")
        p = klt.make_params(kernel=klt.KERNEL_EXACT, **kw)
        out, succ, st = tracker.track(vec["left"], vec["right"], vec["kp1"], vec["kp2"], p)
        assert np.array_equal(out.view(np.uint32), vec[f"{name}_kp2"].view(np.uint32)), name
        assert np.array_equal(succ, vec[f"{name}_succ"]), name
        assert _iters(st, p.levels) == iters and int(st.n_success) == nsucc, name


@pytest.mark.parametrize("kernel", FAST_KERNELS)
def test_fast_kernel_matches_golden_vectors(tracker, golden, kernel):
    meta, vec = golden
    rows, cols = vec["left"].shape
    for name, kw in meta["solver"].items():
        kw = dict(kw)
        iters, _, _src = kw.pop("gn_iters"), kw.pop("n_success"), kw.pop("source")
        if not _supported(kernel, kw):
            continue
        p = klt.make_params(kernel=kernel, **kw)
        out, succ, st = tracker.track(vec["left"], vec["right"], vec["kp1"], vec["kp2"], p)
        assert_parity(out, succ, vec[f"{name}_kp2"], vec[f"{name}_succ"], cols, rows, name)
        assert _iters(st, p.levels) == iters, name  # same convergence decisions


# ------------------------------------------------------------------ BASELINE configs, GPU vs oracle
def _check_case(tracker, oracle, L, R, kp1, kp2, params_kw, kernels, what):
    rows, cols = L.shape
    ref, rs, rst = oracle.track(L, R, kp1, kp2, oracle.make_params(**params_kw), threads=8)
    reports = {}
    for kernel in kernels:
        if not _supported(kernel, params_kw):
            continue
        p = klt.make_params(kernel=kernel, **params_kw)
        out, succ, st = tracker.track(L, R, kp1, kp2, p)
        if kernel == klt.KERNEL_EXACT:
            assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), what
            assert np.array_equal(succ, rs), what
        rep = assert_parity(out, succ, ref, rs, cols, rows, f"{what} kernel={kernel}")
        rep["gn_iters_gpu"], rep["gn_iters_cpu"] = _iters(st, p.levels), _iters(rst, p.levels)
        assert rep["gn_iters_gpu"] == rep["gn_iters_cpu"], (what, rep)
        reports[kernel] = rep
    return reports


@pytest.mark.parametrize("guess", ["same", "noisy"])
def test_config1_single_pair_150_features(tracker, oracle, guess):
    L, R, kp1, kp2, _ = synth.stereo_case(376, 1241, 150, seed=1, guess=guess)
    _check_case(tracker, oracle, L, R, kp1, kp2, dict(levels=4), [klt.KERNEL_EXACT] + FAST_KERNELS, "C1")


def test_config2_sequence_temporal_and_stereo_2000(tracker, oracle):
    for frame in (1, 2):
        P, Cur, kp1, kp2, _ = synth.temporal_case(376, 1241, 2000, seed=2, frame=frame)
        _check_case(tracker, oracle, P, Cur, kp1, kp2, dict(levels=4), [klt.KERNEL_EXACT] + FAST_KERNELS,
                    f"C2 temporal f{frame}")
    L, R, kp1, kp2, _ = synth.stereo_case(376, 1241, 2000, seed=2)
    _check_case(tracker, oracle, L, R, kp1, kp2, dict(levels=4), [klt.KERNEL_EXACT] + FAST_KERNELS, "C2 stereo")


def test_config4_1080p_5_levels_inverse(tracker, oracle):
    L, R, kp1, kp2, _ = synth.stereo_case(1080, 1920, 5000, seed=3)
    _check_case(tracker, oracle, L, R, kp1, kp2, dict(levels=5, inverse=True),
                [klt.KERNEL_EXACT] + FAST_KERNELS, "C4")


def test_config4_shape_forward_all_kernels(tracker, oracle):
    """1920x1080, 5 levels, 5000 features in FORWARD mode: the LANE kernel on the C4 shape (the inverse-mode test above
    exercises EXACT and WARP only), integer and sub-pixel source points."""
    L, R, kp1, kp2, _ = synth.stereo_case(1080, 1920, 5000, seed=3)
    _check_case(tracker, oracle, L, R, kp1, kp2, dict(levels=5), [klt.KERNEL_EXACT] + FAST_KERNELS, "C4 forward")
    rng = np.random.default_rng(33)
    kp1s = (kp1 + rng.uniform(-0.5, 0.5, kp1.shape)).astype(np.float32)
    kp2s = (kp1s + rng.normal(0, 1.5, kp1.shape)).astype(np.float32)
    _check_case(tracker, oracle, L, R, kp1s, kp2s, dict(levels=5), [klt.KERNEL_EXACT] + FAST_KERNELS, "C4 forward sub-pixel")


@pytest.mark.parametrize("n,lo,hi", [(100, -3, 3), (20000, -3, 3), (3000, -5, 5), (3000, -4, 3)])
def test_config5_stress_feature_counts_and_patches(tracker, oracle, n, lo, hi):
    L, R, kp1, kp2, _ = synth.stereo_case(376, 1241, n, seed=4, min_dist=3 if n > 5000 else 5)
    _check_case(tracker, oracle, L, R, kp1, kp2, dict(levels=4, patch_lo=lo, patch_hi=hi), FAST_KERNELS,
                f"C5 n={n} patch={lo}..{hi}")


# ------------------------------------------------------------------ batched path (north star (3))
def _make_batch(B, rows, cols, n, seed0):
    imgs1 = klt.pinned_empty((B, rows, cols), np.uint8)
    imgs2 = klt.pinned_empty((B, rows, cols), np.uint8)
    kp1 = klt.pinned_empty((B, n, 2), np.float32)
    kp2 = klt.pinned_empty((B, n, 2), np.float32)
    for b in range(B):
        L, R, a, g, _ = synth.stereo_case(rows, cols, n, seed=seed0 + b)
        imgs1[b], imgs2[b], kp1[b], kp2[b] = L, R, a, g
    return imgs1, imgs2, kp1, kp2


@pytest.mark.parametrize("kernel", FAST_KERNELS)
def test_config3_batched_pairs_match_oracle_per_pair(tracker, oracle, kernel):
    B, rows, cols, n = 9, 376, 1241, 2000  # >= 8 pairs: exercises the chunked copy/compute overlap
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 1000)
    guess = kp2.copy()
    succ = klt.pinned_empty((B, n), np.uint8)
    batch = tracker.batch(B, rows, cols, n, levels=4)
    st = batch.track(imgs1, imgs2, kp1, kp2, succ, klt.make_params(kernel=kernel))
    assert int(st.n_features) == B * n
    total_iters = np.zeros(4, np.int64)
    for b in range(B):
        ref, rs, rst = oracle.track(imgs1[b], imgs2[b], kp1[b], guess[b], threads=8)
        assert_parity(kp2[b], succ[b], ref, rs, cols, rows, f"C3 pair {b}")
        total_iters += np.array(_iters(rst, 4))
    assert _iters(st, 4) == total_iters.tolist()
    assert int(st.n_success) == int(succ.sum())
    batch.close()


def test_large_batch_fast_kernels_equal_exact_kernel(tracker):
    """More features than resident threads (several features per thread, ring refills with leftovers): every
    output of the fast kernels must be written and match the EXACT kernel, which is bit-identical to the oracle
    (checked above).  Buffers are separate per kernel so an unwritten output cannot hide behind an earlier run."""
    B, rows, cols, n = 64, 188, 620, 2000
    base = [synth.stereo_case(rows, cols, n, seed=3000 + i) for i in range(4)]
    imgs1 = np.stack([base[b % 4][0] for b in range(B)])
    imgs2 = np.stack([base[b % 4][1] for b in range(B)])
    kp1 = np.stack([base[b % 4][2] for b in range(B)])
    kp2 = np.stack([base[b % 4][3] for b in range(B)])
    res = {}
    for kernel in (klt.KERNEL_EXACT, klt.KERNEL_LANE, klt.KERNEL_WARP):
        batch = tracker.batch(B, rows, cols, n, levels=4)
        batch.upload(imgs1, imgs2, kp1, kp2)
        batch.run(klt.make_params(kernel=kernel))
        o, s, st = batch.download()
        res[kernel] = (o, s, _iters(st, 4), int(st.n_success))
        batch.close()
    eo, es, eit, esucc = res[klt.KERNEL_EXACT]
    for kernel in (klt.KERNEL_LANE, klt.KERNEL_WARP):
        o, s, it, ns = res[kernel]
        assert_parity(o, s, eo, es, cols, rows, f"large batch kernel={kernel}")
        assert it == eit and ns == esucc == int(es.sum())


def test_batch_is_idempotent_and_equals_single_calls(tracker):
    """size-independent property: re-running a resident batch gives the same bytes, and each pair of
    the batch equals the single-pair entry point on that pair."""
    B, rows, cols, n = 4, 188, 620, 500
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 2000)
    batch = tracker.batch(B, rows, cols, n, levels=4)
    batch.upload(imgs1, imgs2, kp1, kp2)
    batch.run()
    a, sa, _ = batch.download()
    batch.run()
    b, sb, _ = batch.download()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(sa, sb)
    for i in range(B):
        o, s, _ = tracker.track(imgs1[i], imgs2[i], kp1[i], kp2[i])
        assert np.array_equal(o.view(np.uint32), a[i].view(np.uint32)) and np.array_equal(s, sa[i])
    batch.close()


def test_full_size_config3_size_independent_properties(tracker):
    """BASELINE config C3 at its FULL size (256 pairs 1241x376 x 2000 features, 4 levels), where the CPU oracle is too
    slow to be the checker (bench.py does compare this size with the reference's own code on every run): properties
    that hold at any size.  (a) permuting the pairs permutes the results, byte for byte; (b) permuting the features
    inside every pair likewise; (c) a second run gives the same bytes; (d) a pair tracked against itself from its own
    keypoints stays where it is, exactly, and succeeds; (e) a sample of the pairs equals the bit-exact EXACT kernel."""
    B, rows, cols, n = 256, 376, 1241, 2000
    base = [synth.stereo_case(rows, cols, n, seed=9000 + i) for i in range(8)]
    rng = np.random.default_rng(12)
    imgs1 = klt.pinned_empty((B, rows, cols), np.uint8)
    imgs2 = klt.pinned_empty((B, rows, cols), np.uint8)
    kp1 = klt.pinned_empty((B, n, 2), np.float32)
    kp2 = klt.pinned_empty((B, n, 2), np.float32)
    for b in range(B):
        L, R, a, g, _ = base[b % 8]
        sh = b // 8                       # (every pair distinct: the base pair rolled by whole pixels)
        imgs1[b], imgs2[b] = np.roll(L, sh, axis=1), np.roll(R, sh, axis=1)
        ok = (a[:, 0] + sh < cols - 8)[:, None]                # (the keypoints move with the content where they can)
        kp1[b] = np.where(ok, a + np.float32([sh, 0]), a)
        kp2[b] = np.where(ok, g + np.float32([sh, 0]), g)
        if b % 3 == 0:                    # a third of the pairs with sub-pixel source keypoints
            kp1[b] += rng.uniform(-0.5, 0.5, (n, 2)).astype(np.float32)
            kp2[b] = kp1[b]
    p = klt.make_params(kernel=klt.KERNEL_LANE)
    batch = tracker.batch(B, rows, cols, n, levels=4)

    def run(i1, i2, k1, k2):
        batch.upload(np.ascontiguousarray(i1), np.ascontiguousarray(i2), np.ascontiguousarray(k1), np.ascontiguousarray(k2))
        batch.run(p)
        o, s, st = batch.download()
        return o.copy(), s.copy(), _iters(st, 4), int(st.n_success)

    o0, s0, it0, ns0 = run(imgs1, imgs2, kp1, kp2)
    assert ns0 == int(s0.sum()) and ns0 > 0.3 * B * n
    # (c) same bytes again
    o1, s1, it1, _ = run(imgs1, imgs2, kp1, kp2)
    assert np.array_equal(o0.view(np.uint32), o1.view(np.uint32)) and np.array_equal(s0, s1) and it0 == it1
    # (a) pairs permuted
    perm = rng.permutation(B)
    o2, s2, it2, _ = run(imgs1[perm], imgs2[perm], kp1[perm], kp2[perm])
    assert np.array_equal(o2.view(np.uint32), o0[perm].view(np.uint32)) and np.array_equal(s2, s0[perm]) and it2 == it0
    # (b) features permuted inside every pair
    fperm = rng.permutation(n)
    o3, s3, it3, _ = run(imgs1, imgs2, kp1[:, fperm], kp2[:, fperm])
    assert np.array_equal(o3.view(np.uint32), o0[:, fperm].view(np.uint32)) and np.array_equal(s3, s0[:, fperm]) and it3 == it0
    # (d) identity: img2 = img1, guess = source keypoints
    o4, s4, it4, ns4 = run(imgs1, imgs1, kp1, kp1)
    assert np.array_equal(o4.view(np.uint32), np.asarray(kp1).view(np.uint32)) and ns4 == B * n and bool(s4.all())
    assert it4 == [B * n] * 4             # one pass per level: the update is exactly zero
    batch.close()
    # (e) sixteen of the pairs through the EXACT kernel
    pick = np.sort(rng.choice(B, 16, replace=False))
    chk = tracker.batch(16, rows, cols, n, levels=4)
    chk.upload(np.ascontiguousarray(imgs1[pick]), np.ascontiguousarray(imgs2[pick]), np.ascontiguousarray(kp1[pick]),
               np.ascontiguousarray(kp2[pick]))
    chk.run(klt.make_params(kernel=klt.KERNEL_EXACT))
    eo, es, _ = chk.download()
    assert_parity(o0[pick], s0[pick], eo, es, cols, rows, "C3 full size vs EXACT")
    chk.close()


def _random_problem(seed, allow_non_finite=False):
    """The seeded random problems of tests/test_oracle_vs_ref.py (where the oracle is compared with the reference's own
    code on them): size, texture, keypoints inside / on the border / outside / absurdly far, random guesses."""
    rng = np.random.default_rng(20_000 + seed)
    rows, cols = int(rng.integers(40, 160)), int(rng.integers(40, 220))

    def image(kind):
        if kind == 0:
            return rng.integers(0, 256, size=(rows, cols), dtype=np.uint8)
        if kind == 1:
            yy, xx = np.mgrid[0:rows, 0:cols]
            f = rng.uniform(0.05, 0.6, 4)
            img = 127 + 60 * np.sin(f[0] * xx + f[1] * yy) + 60 * np.cos(f[2] * xx - f[3] * yy)
            return np.clip(img + rng.normal(0, 3, img.shape), 0, 255).astype(np.uint8)
        if kind == 2:
            return np.full((rows, cols), int(rng.integers(0, 256)), np.uint8)
        img = np.zeros((rows, cols), np.uint8)
        img[:, cols // 2:] = 200
        return img

    img1 = image(int(rng.integers(0, 4)))
    img2 = np.roll(img1, (int(rng.integers(-2, 3)), int(rng.integers(-3, 4))), axis=(0, 1)) if rng.random() < 0.7 \
        else image(int(rng.integers(0, 4)))
    n = int(rng.integers(1, 50))
    kp1 = np.stack([rng.uniform(-4, cols + 4, n), rng.uniform(-4, rows + 4, n)], axis=1).astype(np.float32)
    if rng.random() < 0.5:
        kp1 = np.round(kp1)
    kp2 = (kp1 + rng.normal(0, 2.5, kp1.shape)).astype(np.float32)
    special = np.float32([1e7, -1e7, 3e38, -3e38, 65536.5, 0.0, -0.0, 1e-40] + ([np.nan, np.inf, -np.inf] if allow_non_finite else []))
    for _ in range(int(rng.integers(0, 5)) + (3 if allow_non_finite else 0)):
        (kp1 if rng.random() < 0.5 else kp2)[int(rng.integers(0, n)), int(rng.integers(0, 2))] = special[int(rng.integers(0, len(special)))]
    return img1, img2, kp1, kp2, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), (1 if rng.random() < 0.25 else 4)


@pytest.mark.parametrize("kernel", [klt.KERNEL_EXACT] + FAST_KERNELS)
def test_random_problems_match_the_oracle(tracker, oracle, kernel):
    for seed in range(24):
        img1, img2, kp1, kp2, inverse, has_initial, layers = _random_problem(seed)
        kw = dict(levels=layers, inverse=inverse, has_initial=has_initial)
        if not _supported(kernel, kw):
            continue
        ref, rs, rst = oracle.track(img1, img2, kp1, kp2, oracle.make_params(**kw))
        o, s, st = tracker.track(img1, img2, kp1, kp2, klt.make_params(kernel=kernel, **kw))
        if kernel == klt.KERNEL_EXACT:
            assert np.array_equal(o.view(np.uint32), ref.view(np.uint32)) and np.array_equal(s, rs), f"seed {seed}"
        else:
            assert_parity(o, s, ref, rs, img1.shape[1], img1.shape[0], f"random problem {seed} kernel={kernel}")
        assert _iters(st, layers) == _iters(rst, layers), f"seed {seed}"


@pytest.mark.parametrize("kernel", [klt.KERNEL_EXACT] + FAST_KERNELS)
def test_non_finite_keypoints_fail_cleanly(tracker, oracle, kernel):
    """NaN / infinite coordinates: the reference indexes the image with int(NaN) (undefined behaviour, it typically
    crashes; an infinite one becomes NaN through k2 - k1 or is clamped to the border).  Here a NaN feature must simply
    fail -- success = 0 -- and neither kind may disturb the other features of the call, which must equal the oracle's
    results on the call without the non-finite features."""
    for seed in range(8):
        img1, img2, kp1, kp2, inverse, has_initial, layers = _random_problem(100 + seed, allow_non_finite=True)
        kw = dict(levels=layers, inverse=inverse, has_initial=has_initial)
        if not _supported(kernel, kw):
            continue
        # (the guess is read only where has_initial says so: src/algorithm.cpp:47-50)
        nan = np.isnan(kp1).any(axis=1) | (np.isnan(kp2).any(axis=1) & has_initial)
        good = np.isfinite(kp1).all(axis=1) & (np.isfinite(kp2).all(axis=1) | (not has_initial))
        o, s, st = tracker.track(img1, img2, kp1, kp2, klt.make_params(kernel=kernel, **kw))
        assert not s[nan].any(), f"seed {seed}: a feature with NaN coordinates reports success"
        if good.any():
            ref, rs, _ = oracle.track(img1, img2, kp1[good], kp2[good], oracle.make_params(**kw))
            assert_parity(o[good], s[good], ref, rs, img1.shape[1], img1.shape[0], f"non-finite neighbours, seed {seed} kernel={kernel}")


# ------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("kernel", [klt.KERNEL_EXACT] + FAST_KERNELS)
def test_edge_cases(tracker, oracle, kernel):
    rows, cols = 94, 310
    L, R, kp1, kp2, _ = synth.stereo_case(rows, cols, 64, seed=9, min_dist=6)
    # empty list
    out, succ, st = tracker.track(L, R, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32),
                                  klt.make_params(levels=3, kernel=kernel))
    assert out.shape == (0, 2) and succ.shape == (0,) and int(st.n_features) == 0
    # one feature
    o1, s1, _ = tracker.track(L, R, kp1[:1], kp2[:1], klt.make_params(levels=3, kernel=kernel))
    r1, rs1, _ = oracle.track(L, R, kp1[:1], kp2[:1], oracle.make_params(levels=3))
    assert_parity(o1, s1, r1, rs1, cols, rows, "n=1")
    # border / outside / sliver / binade-crossing features, guesses off-image
    ugly = np.array([[0, 0], [1.5, 1.5], [cols - 1, rows - 1], [cols - 0.5, 10], [20, rows - 0.25],
                     [-5, 20], [cols + 3, rows + 3], [127.9999, 63.9999], [255.5, 31.75], [3.9999, 7.9999],
                     [cols - 4.2, rows - 4.1], [2.9, 2.9]], np.float32)
    guess = ugly + np.array([[0.7, -0.4]], np.float32)
    guess[5] = [-30, 20]
    for kw in (dict(levels=3), dict(levels=1), dict(levels=3, inverse=True), dict(levels=3, has_initial=False),
               dict(levels=2, patch_lo=-5, patch_hi=5), dict(levels=3, patch_lo=-4, patch_hi=3)):
        if not _supported(kernel, kw):
            continue
        o, s, _ = tracker.track(L, R, ugly, guess, klt.make_params(kernel=kernel, **kw))
        r, rs, _ = oracle.track(L, R, ugly, guess, oracle.make_params(**kw))
        assert_parity(o, s, r, rs, cols, rows, f"ugly {kw}")
    # flat image: H == 0 -> zero update, success (Eigen LDLT pseudo-inverse)
    flat = np.full((64, 64), 77, np.uint8)
    kp = np.array([[32, 32], [10, 50]], np.float32)
    o, s, st = tracker.track(flat, flat, kp, kp + np.float32(0.25), klt.make_params(levels=2, kernel=kernel))
    assert s.all() and int(st.n_nan) == 0 and np.array_equal(o, kp + np.float32(0.25))


@pytest.mark.parametrize("kernel", [klt.KERNEL_EXACT, klt.KERNEL_WARP, klt.KERNEL_PATCH])
def test_iteration_limits_and_odd_patch_shapes(tracker, oracle, kernel):
    """max_iters 0 / 1 / 3 (the level keeps its guess, one pass, the cap cuts the solve short), a looser and a zero eps,
    1x1, 2x2 (asymmetric), 13x13 patches (six warps of the PATCH kernel), both modes -- kernels that take any
    configuration, against the oracle incl. the pass counts."""
    rows, cols, n = 120, 200, 90
    L, R, kp1, kp2, _ = synth.stereo_case(rows, cols, n, seed=77, min_dist=6, guess="noisy")
    for kw in (dict(max_iters=0), dict(max_iters=1), dict(max_iters=3, inverse=True), dict(eps=0.5), dict(eps=0.0, max_iters=4),
               dict(patch_lo=0, patch_hi=0), dict(patch_lo=-1, patch_hi=0, inverse=True), dict(patch_lo=-6, patch_hi=6),
               dict(patch_lo=-6, patch_hi=6, inverse=True, levels=2), dict(levels=1, has_initial=False)):
        kw = dict(dict(levels=3), **kw)
        ref, rs, rst = oracle.track(L, R, kp1, kp2, oracle.make_params(**kw))
        o, s, st = tracker.track(L, R, kp1, kp2, klt.make_params(kernel=kernel, **kw))
        if kernel == klt.KERNEL_EXACT:
            assert np.array_equal(o.view(np.uint32), ref.view(np.uint32)) and np.array_equal(s, rs), kw
        else:
            assert_parity(o, s, ref, rs, cols, rows, f"{kw} kernel={kernel}")
        assert _iters(st, kw["levels"]) == _iters(rst, kw["levels"]), kw


@pytest.mark.parametrize("lo,hi", [(-3, 3), (-4, 3), (-5, 5)])
def test_subpixel_keypoints_and_deferred_features(tracker, oracle, lo, hi):
    """Tracked (sub-pixel) source points, as Frontend::TrackLastFrame feeds them: float(kx+c) is sometimes
    rounded near powers of two.  The LANE kernel splits such a patch into coordinate families (one extra trip
    per pass); only patches straddling several powers of two go to the exact warp kernel.  All three compiled
    patch instances."""
    rows, cols, n = 376, 1241, 4000
    L, R, kp1, kp2, _ = synth.stereo_case(rows, cols, n, seed=21)
    rng = np.random.default_rng(5)
    kp1 = (kp1 + rng.uniform(-0.5, 0.5, kp1.shape)).astype(np.float32)
    kp2 = (kp1 + rng.normal(0, 1.0, kp1.shape)).astype(np.float32)
    ref, rs, rst = oracle.track(L, R, kp1, kp2, oracle.make_params(patch_lo=lo, patch_hi=hi), threads=8)
    for kernel in FAST_KERNELS:
        out, succ, st = tracker.track(L, R, kp1, kp2, klt.make_params(kernel=kernel, patch_lo=lo, patch_hi=hi))
        assert_parity(out, succ, ref, rs, cols, rows, f"subpixel kernel={kernel}")
        assert _iters(st, 4) == _iters(rst, 4)
        if kernel == klt.KERNEL_LANE:
            assert int(st.n_deferred) < n // 50, int(st.n_deferred)


def test_sequence_mode_image_handles_equal_pairwise_calls(tracker, oracle):
    """Frontend::Track does temporal (last-left -> cur-left) and stereo (cur-left -> cur-right)
    tracking per frame; with image handles every image is uploaded and pyramided once, and the results are the
    same bytes as the pairwise entry point (which the oracle parity tests cover)."""
    rows, cols, n = 376, 1241, 2000
    lefts = [synth.next_frame(rows, cols, 2, f)[0] for f in range(3)]
    rights = [np.ascontiguousarray(np.roll(img, -7, axis=1)) for img in lefts]
    h_left = [tracker.image(rows, cols, 4).upload(img) for img in lefts]
    h_right = [tracker.image(rows, cols, 4).upload(img) for img in rights]
    kp = synth.detect_features(lefts[0], n, min_dist=5, seed=3)
    for f in (1, 2):
        for kernel in (klt.KERNEL_AUTO, klt.KERNEL_WARP):
            p = klt.make_params(kernel=kernel)
            a, sa, st_a = tracker.track_images(h_left[f - 1], h_left[f], kp, kp, p)        # temporal
            b, sb, st_b = tracker.track(lefts[f - 1], lefts[f], kp, kp, p)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(sa, sb)
            assert _iters(st_a, 4) == _iters(st_b, 4)
            c, sc, _ = tracker.track_images(h_left[f], h_right[f], a, a, p)               # stereo on tracked points
            d, sd, _ = tracker.track(lefts[f], rights[f], a, a, p)
            assert np.array_equal(c.view(np.uint32), d.view(np.uint32)) and np.array_equal(sc, sd)
        ref, rs, _ = oracle.track(lefts[f - 1], lefts[f], kp, kp, threads=8)
        assert_parity(a, sa, ref, rs, cols, rows, f"sequence frame {f}")
        kp = a  # tracked (sub-pixel) points feed the next frame, like the reference's frontend


def test_lane_kernel_rejects_unsupported_configurations(tracker):
    from lego_slam_b200 import _lib
    img = np.zeros((64, 64), np.uint8)
    kp = np.full((1, 2), 30, np.float32)
    with pytest.raises(_lib.KltError):   # inverse mode is compiled for the reference's 7x7 patch only
        tracker.track(img, img, kp, kp, klt.make_params(levels=2, inverse=True, patch_lo=-4, patch_hi=3, kernel=klt.KERNEL_LANE))
    with pytest.raises(_lib.KltError):
        tracker.track(img, img, kp, kp, klt.make_params(levels=2, patch_lo=-6, patch_hi=6, kernel=klt.KERNEL_LANE))
    tracker.track(img, img, kp, kp, klt.make_params(levels=2, inverse=True, kernel=klt.KERNEL_LANE))   # supported


@pytest.mark.gpu
@pytest.mark.parametrize("lo,hi", [(-4, 3), (-5, 5)])
def test_lane_kernel_other_patches_on_a_large_batch_equal_exact_kernel(tracker, lo, hi):
    """The 8x8 (-4..3: the patch BASELINE.json's metric text names) and 11x11 (-5..5: the stress configuration)
    instances of the LANE kernel against the bit-exact EXACT kernel on a batch large enough to fill the persistent
    kernel: flags equal, positions within the contract, identical iteration counts."""
    B, rows, cols, n = 12, 376, 1241, 2000
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 2000)
    res = {}
    for name, k in (("exact", klt.KERNEL_EXACT), ("lane", klt.KERNEL_LANE), ("auto", klt.KERNEL_AUTO)):
        batch = tracker.batch(B, rows, cols, n, levels=4)
        batch.upload(imgs1, imgs2, kp1, kp2)
        batch.run(klt.make_params(patch_lo=lo, patch_hi=hi, kernel=k))
        o, s_, st = batch.download()
        res[name] = (o.copy(), s_.copy(), _iters(st, 4))
    eo, es, ei = res["exact"]
    for name in ("lane", "auto"):
        o, s_, it = res[name]
        assert np.array_equal(s_, es), name
        assert np.abs(o.astype(np.float64) - eo).max() <= 1e-3, name
        assert it == ei, name
    assert (res["lane"][0].view(np.uint32) == eo.view(np.uint32)).all(axis=2).mean() >= 0.999


def test_bad_arguments_are_rejected(tracker):
    from lego_slam_b200 import _lib
    img = np.zeros((32, 32), np.uint8)
    kp = np.zeros((1, 2), np.float32)
    with pytest.raises(_lib.KltError):
        tracker.track(img, img, kp, kp, klt.make_params(levels=9))
    with pytest.raises(_lib.KltError):
        tracker.track(img, img, kp, kp, klt.make_params(levels=7))  # level would be empty
    with pytest.raises(_lib.KltError):
        tracker.track(img, img, kp, kp, klt.make_params(levels=1, patch_lo=-8, patch_hi=8))


# ------------------------------------------------------------------ image ingest
@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,pad", [(376, 1241, 0), (375, 1242, 0), (751, 2483, 13), (3, 5, 0), (1080, 1920, 0)])
def test_ingest_half_nearest_equals_oracle_and_cv2(tracker, rows, cols, pad):
    """Dataset::NextFrame: cv::resize(img, out, cv::Size(), 0.5, 0.5, cv::INTER_NEAREST) (src/dataset.cpp:75-77).
    OpenCV is third party; the pin is Python cv2 in this image (same call)."""
    from oracle import ingest_np
    big = np.random.default_rng(rows * 7 + cols).integers(0, 256, size=(rows, cols + pad), dtype=np.uint8)
    full = big[:, :cols]
    ref = ingest_np.downscale_half_nearest(full)      # (pinned against cv2 by tests/test_oracle_pyramid.py)
    got = tracker.downscale_half(full)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    try:
        import cv2
        assert np.array_equal(got, cv2.resize(np.ascontiguousarray(full), None, fx=0.5, fy=0.5, interpolation=cv2.INTER_NEAREST))
    except ImportError:
        pass


@pytest.mark.gpu
def test_fullres_upload_tracks_like_the_halved_images(tracker, oracle):
    """Frame ingest + sequence mode: handles fed with full-resolution frames give the same bytes as tracking the
    host-halved images (and hence the oracle)."""
    cv2 = pytest.importorskip("cv2")
    L, R, kp1, kp2, _ = synth.stereo_case(376, 1241, 300, seed=77)
    halve = lambda a: cv2.resize(a, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_NEAREST)
    l2, r2 = halve(L), halve(R)
    k1 = (kp1 * 0.5).astype(np.float32)
    k1 = k1[(k1[:, 0] > 8) & (k1[:, 0] < l2.shape[1] - 8) & (k1[:, 1] > 8) & (k1[:, 1] < l2.shape[0] - 8)]
    a = tracker.image(l2.shape[0], l2.shape[1]).upload_fullres(L)
    b = tracker.image(l2.shape[0], l2.shape[1]).upload_fullres(R)
    out, succ, st = tracker.track_images(a, b, k1, k1)
    ref, rs, _ = oracle.track(l2, r2, k1, k1)
    assert_parity(out, succ, ref, rs, l2.shape[1], l2.shape[0], "fullres ingest")
    out2, succ2, _ = tracker.track(l2, r2, k1, k1)
    assert np.array_equal(out.view(np.uint32), out2.view(np.uint32)) and np.array_equal(succ, succ2)


@pytest.mark.gpu
def test_ingest_reproduces_the_committed_cv2_hashes(tracker):
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next_rows.json")) as f:
        meta = json.load(f)
    for name, c in meta["ingest"].items():
        img = np.random.default_rng(c["seed"]).integers(0, 256, size=(c["rows"], c["cols"]), dtype=np.uint8)
        out = tracker.downscale_half(img)
        assert list(out.shape) == c["out_shape"], name
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == c["sha256"], name


# ------------------------------------------------------------------ handle lifetime
def test_handles_survive_their_tracker_in_any_destruction_order():
    """Image and Batch handles point into their context: destroying the context first must only defer its tear-down
    (include/lego_klt.h, ownership note)."""
    import gc
    L, R, kp1, kp2, _ = synth.stereo_case(94, 310, 50, seed=8, min_dist=8)
    trk = klt.Tracker(0)
    im = trk.image(94, 310, 3).upload(L)
    batch = trk.batch(2, 94, 310, 50, levels=3)
    trk.close()            # lego_klt_destroy with two live handles
    del trk
    gc.collect()
    batch.close()          # ... which are destroyed afterwards, the last one tears the context down
    im.close()
    # temporaries: the tracker object would be collected before the image without the back-reference
    im2 = klt.Tracker(0).image(94, 310, 3)
    gc.collect()
    im2.upload(R)
    im2.close()


# ------------------------------------------------------------------ multi-rank on the GPU
def _gpu_rank(rank, world, port, out_dir, B, rows, cols, n):
    import os
    import torch
    import torch.distributed as dist
    from lego_slam_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 4000)
    lo, hi = sharding.shard_range(B, rank, world)
    trk = klt.Tracker(dev)
    batch = trk.batch(hi - lo, rows, cols, n, levels=4)
    succ = klt.pinned_empty((hi - lo, n), np.uint8)
    io = klt.pinned_empty((hi - lo, n, 2), np.float32)
    np.copyto(io, kp2[lo:hi])
    batch.track(np.ascontiguousarray(imgs1[lo:hi]), np.ascontiguousarray(imgs2[lo:hi]),
                np.ascontiguousarray(kp1[lo:hi]), io, succ, klt.make_params(kernel=klt.KERNEL_LANE))
    kp_full, su_full = sharding.gather_results(torch.from_numpy(io.copy()), torch.from_numpy(succ.copy()), B)
    if rank == 0:
        np.save(os.path.join(out_dir, "kp.npy"), kp_full.numpy())
        np.save(os.path.join(out_dir, "su.npy"), su_full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharded_gpu_result_equals_single_gpu_bytes(tracker, tmp_path):
    """Block partition over 2 ranks (each its own process and context; on a 1-GPU box both use cuda:0), final gather:
    byte-for-byte what ONE context computes for the whole batch."""
    import socket
    import torch.multiprocessing as mp
    B, rows, cols, n = 5, 188, 620, 1500     # uneven shards (3 + 2); the LANE kernel, explicitly
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gpu_rank, args=(2, port, str(tmp_path), B, rows, cols, n), nprocs=2, join=True)
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 4000)
    succ = klt.pinned_empty((B, n), np.uint8)
    batch = tracker.batch(B, rows, cols, n, levels=4)
    batch.track(imgs1, imgs2, kp1, kp2, succ, klt.make_params(kernel=klt.KERNEL_LANE))
    assert np.array_equal(np.load(tmp_path / "kp.npy").view(np.uint32), kp2.view(np.uint32))
    assert np.array_equal(np.load(tmp_path / "su.npy"), succ)
    batch.close()


# ------------------------------------------------------------------ ragged batches, device lists
@pytest.mark.parametrize("kernel", [klt.KERNEL_EXACT, klt.KERNEL_WARP, klt.KERNEL_LANE])
def test_ragged_batch_tracks_only_the_first_counts_features_of_each_pair(tracker, kernel):
    B, rows, cols, n = 9, 188, 620, 1200
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 5000)
    guess = kp2.copy()
    counts = np.array([n, 0, 1, 37, n - 1, 600, n, 999, 5], np.int32)
    batch = tracker.batch(B, rows, cols, n, levels=4)
    full = tracker.batch(B, rows, cols, n, levels=4)
    succ_f = klt.pinned_empty((B, n), np.uint8)
    kp_f = klt.pinned_empty((B, n, 2), np.float32)
    np.copyto(kp_f, guess)
    full.track(imgs1, imgs2, kp1, kp_f, succ_f, klt.make_params(kernel=kernel))
    batch.set_feature_counts(counts)
    succ = klt.pinned_empty((B, n), np.uint8)
    succ[:] = 7
    st = batch.track(imgs1, imgs2, kp1, kp2, succ, klt.make_params(kernel=kernel))
    assert int(st.n_features) == int(counts.sum())
    for b in range(B):
        c = counts[b]
        assert np.array_equal(kp2[b, :c].view(np.uint32), kp_f[b, :c].view(np.uint32)), b
        assert np.array_equal(succ[b, :c], succ_f[b, :c]), b
        assert np.array_equal(kp2[b, c:].view(np.uint32), guess[b, c:].view(np.uint32)), b   # untouched slots
        assert not succ[b, c:].any(), b
    assert int(st.n_success) == int(sum(succ_f[b, :counts[b]].sum() for b in range(B)))
    batch.set_feature_counts(None)       # back to a full batch
    np.copyto(kp2, guess)
    batch.track(imgs1, imgs2, kp1, kp2, succ, klt.make_params(kernel=kernel))
    assert np.array_equal(kp2.view(np.uint32), kp_f.view(np.uint32)) and np.array_equal(succ, succ_f)
    batch.close()
    full.close()


def test_multi_device_entry_equals_single_context_bytes(tracker):
    """lego_klt_multi_track over a device list (two contexts; both on cuda:0 when the box has one GPU)."""
    import torch
    B, rows, cols, n = 11, 188, 620, 1500
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 6000)
    guess = kp2.copy()
    succ1 = klt.pinned_empty((B, n), np.uint8)
    one = tracker.batch(B, rows, cols, n, levels=4)
    st1 = one.track(imgs1, imgs2, kp1, kp2, succ1, klt.make_params(kernel=klt.KERNEL_LANE))
    ref = kp2.copy()
    ndev = torch.cuda.device_count()
    multi = klt.MultiTracker([0, 1 % ndev, 0], B, rows, cols, n, levels=4)
    assert [s[1:] for s in multi.shards()] == [(0, 4), (4, 4), (8, 3)]
    np.copyto(kp2, guess)
    succ2 = klt.pinned_empty((B, n), np.uint8)
    st2 = multi.track(imgs1, imgs2, kp1, kp2, succ2, klt.make_params(kernel=klt.KERNEL_LANE))
    assert np.array_equal(kp2.view(np.uint32), ref.view(np.uint32)) and np.array_equal(succ1, succ2)
    assert list(st1.gn_iters) == list(st2.gn_iters) and int(st1.n_success) == int(st2.n_success)
    assert int(st2.n_features) == B * n
    multi.close()
    one.close()


def test_multi_device_dynamic_schedule_equals_single_context_bytes(tracker):
    """lego_klt_multi_set_schedule: devices pull blocks of pairs from one counter (two in flight per device); the batch
    size is not a multiple of the block, one pass is ragged: bytes equal a single context's, every pair tracked once."""
    import torch
    B, rows, cols, n = 23, 188, 620, 1500
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 6100)
    guess = kp2.copy()
    ndev = torch.cuda.device_count()
    counts = np.random.default_rng(3).integers(0, n + 1, B).astype(np.int32)
    counts[5] = 0
    for ragged in (False, True):
        one = tracker.batch(B, rows, cols, n, levels=4)
        one.set_feature_counts(counts if ragged else None)
        succ1 = klt.pinned_empty((B, n), np.uint8)
        succ1[:] = 7
        np.copyto(kp2, guess)
        st1 = one.track(imgs1, imgs2, kp1, kp2, succ1, klt.make_params(kernel=klt.KERNEL_LANE))
        ref = kp2.copy()
        one.close()
        multi = klt.MultiTracker([0, 1 % ndev, 0], B, rows, cols, n, levels=4)
        multi.set_feature_counts(counts if ragged else None)
        multi.set_schedule(4)
        for _ in range(2):   # (the second call re-uses the block objects)
            np.copyto(kp2, guess)
            succ2 = klt.pinned_empty((B, n), np.uint8)
            succ2[:] = 7
            st2 = multi.track(imgs1, imgs2, kp1, kp2, succ2, klt.make_params(kernel=klt.KERNEL_LANE))
            used = np.broadcast_to(np.arange(n)[None, :] < (counts[:, None] if ragged else n), (B, n))
            assert np.array_equal(kp2.view(np.uint32)[used], ref.view(np.uint32)[used])
            assert np.array_equal(succ1[used], succ2[used])
            assert list(st1.gn_iters) == list(st2.gn_iters) and int(st1.n_success) == int(st2.n_success)
            assert int(st2.n_features) == (int(counts.sum()) if ragged else B * n)
            dist = multi.last_distribution()
            assert len(dist) == 3 and sum(dist) == B
        multi.set_schedule(0)    # back to the static blocks
        np.copyto(kp2, guess)
        multi.track(imgs1, imgs2, kp1, kp2, succ2, klt.make_params(kernel=klt.KERNEL_LANE))
        assert np.array_equal(kp2.view(np.uint32)[used], ref.view(np.uint32)[used])
        assert multi.last_distribution() == [8, 8, 7]
        multi.close()


def test_kernel_launch_counter_counts_this_librarys_launches(tracker):
    L, R, kp1, kp2, _ = synth.stereo_case(94, 310, 50, seed=8, min_dist=8)
    before = klt.kernel_launches()
    tracker.track(L, R, kp1, kp2, klt.make_params(levels=3, kernel=klt.KERNEL_EXACT))
    # 2 ingest + level 0->1 + band kernel + exact solver
    assert klt.kernel_launches() - before == 5


def test_pipeline_chunk_count_does_not_change_the_bytes(tracker):
    """lego_klt_track_batched with 1 / 3 / 16 chunks of H2D - compute - D2H overlap, integer and sub-pixel keypoints (the
    family instance runs on the side stream and is joined per chunk): identical results."""
    B, rows, cols, n = 33, 188, 620, 1000
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 7000)
    rng = np.random.default_rng(5)
    for jitter in (False, True):
        k1 = kp1.copy()
        if jitter:
            k1 += rng.uniform(-0.5, 0.5, k1.shape).astype(np.float32)
        k1p = klt.pinned_empty(k1.shape, np.float32)
        np.copyto(k1p, k1)
        ref = None
        for chunks in (1, 3, 16, 0):
            batch = tracker.batch(B, rows, cols, n, levels=4)
            batch.set_pipeline_chunks(chunks)
            io = klt.pinned_empty((B, n, 2), np.float32)
            np.copyto(io, k1)
            succ = klt.pinned_empty((B, n), np.uint8)
            st = batch.track(imgs1, imgs2, k1p, io, succ, klt.make_params())
            cur = (io.copy(), succ.copy(), list(st.gn_iters), int(st.n_success))
            if ref is None:
                ref = cur
            else:
                assert np.array_equal(cur[0].view(np.uint32), ref[0].view(np.uint32)) and np.array_equal(cur[1], ref[1])
                assert cur[2] == ref[2] and cur[3] == ref[3]
            batch.close()


@pytest.mark.parametrize("n", [150, 2000, 6000])
def test_fused_frame_call_equals_two_calls(tracker, oracle, n):
    """lego_klt_track_frame (temporal track + device-chained stereo match of the kept features, one synchronisation)
    against lego_klt_track_images twice, and against the oracle run the same way; n = 6000 takes the LANE kernel."""
    P, Cur, kt, _, _ = synth.temporal_case(376, 1241, n, seed=2, frame=1)
    _, Rt, _ = synth.stereo_pair(376, 1241, 2)
    kt[:3] = [[1.0, 1.0], [1239.5, 374.5], [-3.0, 10.0]]       # some that fail / leave the image
    hp, hc, hr = (tracker.image(376, 1241, 4).upload(x) for x in (P, Cur, Rt))
    cur, st_, right, ss_, a, b = tracker.track_frame(hp, hc, hr, kt, kt, want_stats=True)
    cur2, st2, _ = tracker.track_images(hp, hc, kt, kt)
    assert np.array_equal(cur.view(np.uint32), cur2.view(np.uint32)) and np.array_equal(st_, st2)
    keep = st2.astype(bool)
    r2, s2, stat2 = tracker.track_images(hc, hr, cur2[keep], cur2[keep])
    assert np.array_equal(right[keep].view(np.uint32), r2.view(np.uint32)) and np.array_equal(ss_[keep], s2)
    assert not ss_[~keep].any() and np.array_equal(right[~keep].view(np.uint32), cur[~keep].view(np.uint32))
    assert int(b.n_success) == int(s2.sum()) and list(b.gn_iters) == list(stat2.gn_iters)
    ref_c, ref_s, _ = oracle.track(P, Cur, kt, kt, threads=8)
    assert np.array_equal(ref_s, st_) and np.abs(ref_c - cur).max() <= 1e-3


def test_begin_end_halves_with_two_calls_in_flight_equal_the_synchronous_call(tracker):
    """lego_klt_track_batched_begin / _end on two batch objects of two contexts, overlapping, against the one-call form."""
    B, rows, cols, n = 40, 188, 620, 1200
    imgs1, imgs2, kp1, kp2 = _make_batch(B, rows, cols, n, 8000)
    ref = klt.pinned_empty((B, n, 2), np.float32)
    np.copyto(ref, kp2)
    ref_s = klt.pinned_empty((B, n), np.uint8)
    one = tracker.batch(B, rows, cols, n, levels=4)
    st0 = one.track(imgs1, imgs2, kp1, ref, ref_s, klt.make_params())
    trk2 = klt.Tracker(0)
    objs = [one, trk2.batch(B, rows, cols, n, levels=4)]
    bufs = [(klt.pinned_empty((B, n, 2), np.float32), klt.pinned_empty((B, n), np.uint8)) for _ in range(2)]
    for rounds in range(3):
        for j in range(2):
            np.copyto(bufs[j][0], kp2)
            objs[j].track_begin(imgs1, imgs2, kp1, bufs[j][0], bufs[j][1], klt.make_params())
        for j in range(2):
            st = objs[j].track_end()
            assert np.array_equal(bufs[j][0].view(np.uint32), ref.view(np.uint32)) and np.array_equal(bufs[j][1], ref_s)
            assert list(st.gn_iters) == list(st0.gn_iters) and int(st.n_success) == int(st0.n_success)
    with pytest.raises(Exception):
        objs[0].track_end()          # nothing in flight
    objs[1].close()
    one.close()
