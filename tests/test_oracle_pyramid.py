"""Pins the oracle's pyramid (cv::resize restatement, src/algorithm.cpp:147-150) to OpenCV itself."""
import hashlib

import numpy as np
import pytest


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_pyramid_matches_committed_cv2_hashes(oracle, golden):
    meta, _ = golden
    for name, c in meta["pyramid"].items():
        img = np.random.default_rng(c["seed"]).integers(0, 256, size=(c["rows"], c["cols"]), dtype=np.uint8)
        pyr = oracle.build_pyramid(img, c["levels"])
        for l, h in enumerate(c["levels_sha"], start=1):
            assert list(pyr[l].shape) == h["shape"], (name, l)
            assert _sha(pyr[l]) == h["sha256"], (name, l)


@pytest.mark.parametrize("rows,cols,levels", [(376, 1241, 4), (188, 620, 4), (1080, 1920, 5), (217, 333, 3),
                                              (95, 157, 4), (64, 64, 5), (9, 11, 2)])
def test_pyramid_bit_exact_vs_cv2(oracle, rows, cols, levels):
    cv2 = pytest.importorskip("cv2")
    img = np.random.default_rng(rows * 7 + cols).integers(0, 256, size=(rows, cols), dtype=np.uint8)
    pyr = oracle.build_pyramid(img, levels)
    cur = img
    for l in range(1, levels):
        cur = cv2.resize(cur, (int(cur.shape[1] * 0.5), int(cur.shape[0] * 0.5)))
        assert np.array_equal(cur, pyr[l]), f"level {l}: {(cur != pyr[l]).sum()} pixels differ"


def test_exact_halving_is_rounded_box_mean(oracle):
    img = np.random.default_rng(5).integers(0, 256, size=(94, 310), dtype=np.uint8)
    half = oracle.resize_half(img)
    a = img.astype(np.int32)
    box = (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2
    assert np.array_equal(half, box.astype(np.uint8))


def test_resize_respects_row_step(oracle):
    big = np.random.default_rng(6).integers(0, 256, size=(60, 100), dtype=np.uint8)
    view = big[:, :77]  # step 100, cols 77
    assert np.array_equal(oracle.resize_half(view), oracle.resize_half(np.ascontiguousarray(view)))


def test_empty_level_is_an_error(oracle):
    img = np.zeros((3, 3), np.uint8)
    with pytest.raises(RuntimeError):
        oracle.build_pyramid(img, 4)


@pytest.mark.parametrize("rows,cols", [(376, 1241), (375, 1242), (751, 2483), (3, 5), (1080, 1920), (2, 2), (2, 3)])  # (an axis of 1 pixel halves to 0: OpenCV asserts, the C ABI returns BAD_ARG)
def test_ingest_oracle_equals_cv2_nearest_halving(rows, cols):
    """Dataset::NextFrame (src/dataset.cpp:75-77): the numpy restatement against the only OpenCV in this image."""
    cv2 = pytest.importorskip("cv2")
    from oracle import ingest_np
    img = np.random.default_rng(rows * 31 + cols).integers(0, 256, size=(rows, cols), dtype=np.uint8)
    ref = cv2.resize(img, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_NEAREST)
    got = ingest_np.downscale_half_nearest(img)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_ingest_oracle_reproduces_the_committed_cv2_hashes():
    """tests/golden/next_rows.json (tools/make_golden_next.py): SHA-256 of cv2's own INTER_NEAREST halving."""
    import json
    import os
    from oracle import ingest_np
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next_rows.json")) as f:
        meta = json.load(f)
    for name, c in meta["ingest"].items():
        img = np.random.default_rng(c["seed"]).integers(0, 256, size=(c["rows"], c["cols"]), dtype=np.uint8)
        out = ingest_np.downscale_half_nearest(img)
        assert list(out.shape) == c["out_shape"] and _sha(out) == c["sha256"], name
