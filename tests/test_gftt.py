"""Feature detection: Frontend::DetectFeatures = cv::goodFeaturesToTrack(img, n, 0.01, 20, mask)
(/root/reference src/frontend_g2o.cpp:16,279-297).  OpenCV is third party; the parity definition is tolerance-aware
(oracle/gftt_np.py): eigenvalue map within EIG_TOL_REL of the structure-tensor trace, corner lists equal except for
decisions that are ambiguous at that tolerance (reported, few).  The CUDA path must equal the numpy oracle EXACTLY
(same fp32 operations in the same order)."""
import numpy as np
import pytest

from lego_slam_b200 import synth
from oracle import gftt_np as g


def _images():
    out = []
    for seed, (rows, cols) in ((1, (376, 1241)), (2, (188, 620)), (5, (97, 131))):
        L, _, _ = synth.stereo_pair(rows, cols, seed)
        out.append(L)
    rng = np.random.default_rng(3)
    out.append(rng.integers(0, 256, size=(64, 80), dtype=np.uint8))          # white noise: many near-ties
    flat = np.full((48, 64), 90, np.uint8)
    flat[20:30, 25:40] = 200                                                  # one bright rectangle: plateaus, exact ties
    out.append(flat)
    return out


# ---------------------------------------------------------------- CPU: the oracle against OpenCV itself
def test_oracle_eigenvalue_map_matches_cv2_within_tolerance():
    cv2 = pytest.importorskip("cv2")
    for img in _images():
        e_cv = cv2.cornerMinEigenVal(img, 3, ksize=3)
        e, tr = g.corner_min_eigen_val(img)
        d = np.abs(e.astype(np.float64) - e_cv)
        assert (d <= g.EIG_TOL_REL * np.maximum(tr, 1e-30) + 1e-30).all(), float((d / np.maximum(tr, 1e-30)).max())


@pytest.mark.parametrize("n,md", [(150, 20.0), (2000, 5.0), (500, 0.0), (50, 33.5)])
def test_oracle_corner_lists_match_cv2(n, md):
    cv2 = pytest.importorskip("cv2")
    total_only = 0
    for img in _images():
        e_cv = cv2.cornerMinEigenVal(img, 3, ksize=3)
        _, tr = g.corner_min_eigen_val(img)
        pts, _ = g.good_features_to_track(img, n, 0.01, md)
        cvp = cv2.goodFeaturesToTrack(img, n, 0.01, md)
        cvp = np.zeros((0, 2), np.float32) if cvp is None else cvp.reshape(-1, 2)
        rep = g.compare_corner_lists(pts, cvp, e_cv, tr, max(md, 1.0))
        assert not rep["unexplained"], rep
        assert abs(rep["n_ours"] - rep["n_theirs"]) <= max(2, rep["n_only_one_list"])
        total_only += rep["n_only_one_list"]
    assert total_only <= 10   # ambiguous decisions are rare


def test_oracle_mask_and_exclusion_rectangles_match_cv2():
    cv2 = pytest.importorskip("cv2")
    img = _images()[0]
    rows, cols = img.shape
    rng = np.random.default_rng(8)
    existing = np.stack([rng.uniform(-5, cols + 5, 40), rng.uniform(-5, rows + 5, 40)], axis=1).astype(np.float32)
    existing[:4] = [[10.5, 10.5], [11.5, 20.5], [0.0, 0.0], [cols - 1.0, rows - 1.0]]   # .5: Point2f -> Point rounds half to even
    mask_cv = np.full((rows, cols), 255, np.uint8)
    for x, y in existing:
        # cv::rectangle(mask, pt - Point2f(10, 10), pt + Point2f(10, 10), 0, CV_FILLED) -- the Point2f arguments convert
        # to cv::Point by cvRound
        p1 = (int(np.rint(np.float32(x - 10))), int(np.rint(np.float32(y - 10))))
        p2 = (int(np.rint(np.float32(x + 10))), int(np.rint(np.float32(y + 10))))
        cv2.rectangle(mask_cv, p1, p2, 0, -1)
    mask = g.exclusion_mask(rows, cols, existing, 10.0)
    assert np.array_equal(mask, mask_cv)
    pts, _ = g.good_features_to_track(img, 150, 0.01, 20.0, mask)
    cvp = cv2.goodFeaturesToTrack(img, 150, 0.01, 20.0, mask=mask_cv).reshape(-1, 2)
    e_cv = cv2.cornerMinEigenVal(img, 3, ksize=3)
    _, tr = g.corner_min_eigen_val(img)
    rep = g.compare_corner_lists(pts, cvp, e_cv, tr, 20.0)
    assert not rep["unexplained"] and rep["n_only_one_list"] <= 2, rep
    assert all(mask[int(y), int(x)] for x, y in pts)


# ---------------------------------------------------------------- GPU: the CUDA path against the oracle (exact) and cv2
@pytest.mark.gpu
def test_gpu_eigenvalue_map_equals_oracle_bitwise(tracker):
    for img in _images():
        tracker.detect_features(img, 10)
        e_gpu = tracker.debug_read_eig(*img.shape)
        e, _ = g.corner_min_eigen_val(img)
        assert np.array_equal(e_gpu.view(np.uint32), e.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("n,md", [(150, 20.0), (2000, 5.0), (500, 0.0), (50, 33.5), (100000, 1.0)])
def test_gpu_corner_list_equals_oracle_exactly(tracker, n, md):
    for img in _images():
        pts, sc = tracker.detect_features(img, n, 0.01, md)
        ref, rsc = g.good_features_to_track(img, n, 0.01, md)
        assert np.array_equal(pts, ref) and np.array_equal(sc.view(np.uint32), rsc.view(np.uint32))


@pytest.mark.gpu
def test_gpu_detection_with_mask_exclusion_list_and_image_handle(tracker):
    img = _images()[0]
    rows, cols = img.shape
    rng = np.random.default_rng(9)
    existing = np.stack([rng.uniform(-5, cols + 5, 60), rng.uniform(-5, rows + 5, 60)], axis=1).astype(np.float32)
    existing[:2] = [[10.5, 10.5], [11.5, 20.5]]
    mask = g.exclusion_mask(rows, cols, existing, 10.0)
    ref, _ = g.good_features_to_track(img, 150, 0.01, 20.0, mask)
    a, _ = tracker.detect_features(img, 150, 0.01, 20.0, mask=mask)                       # cv::Mat mask
    b, _ = tracker.detect_features(img, 150, 0.01, 20.0, exclude=existing)                # the reference's rectangles, on the device
    h = tracker.image(rows, cols, 4).upload(img)
    c, _ = tracker.detect_features(h, 150, 0.01, 20.0, exclude=existing)                  # image already in HBM
    assert np.array_equal(a, ref) and np.array_equal(b, ref) and np.array_equal(c, ref)
    padded = np.zeros((rows, cols + 23), np.uint8)                                        # cv::Mat with step > cols
    padded[:, :cols] = img
    d, _ = tracker.detect_features(padded[:, :cols], 150, 0.01, 20.0, mask=mask)
    assert np.array_equal(d, ref)
    e, _ = tracker.detect_features(img, 150, 0.01, 20.0, mask=np.zeros_like(img))         # nothing allowed: no corners
    assert e.shape == (0, 2)


@pytest.mark.gpu
def test_gpu_detection_matches_cv2_within_the_stated_tolerance(tracker):
    cv2 = pytest.importorskip("cv2")
    for img in _images()[:3]:
        for n, md in ((150, 20.0), (2000, 5.0)):
            pts, _ = tracker.detect_features(img, n, 0.01, md)
            cvp = cv2.goodFeaturesToTrack(img, n, 0.01, md).reshape(-1, 2)
            e_cv = cv2.cornerMinEigenVal(img, 3, ksize=3)
            _, tr = g.corner_min_eigen_val(img)
            rep = g.compare_corner_lists(pts, cvp, e_cv, tr, md)
            assert not rep["unexplained"] and rep["n_only_one_list"] <= 4, rep


@pytest.mark.gpu
@pytest.mark.parametrize("n,md,exclude", [(150, 20.0, False), (150, 20.0, True), (800, 5.0, True), (60, 0.0, False)])
def test_gpu_batched_detection_equals_the_single_image_calls(tracker, n, md, exclude):
    """lego_klt_batch_detect_features (every image of a batch per launch, more images than one workspace pass holds, both
    image sets, the exclusion mask from the pair's own source keypoints, a ragged batch) against the single-image entry
    point on each image -- which equals the numpy oracle exactly (above)."""
    import lego_slam_b200 as klt
    B, rows, cols, nk = 37, 120, 200, 40
    rng = np.random.default_rng(21)
    base = _images()[:2]
    imgs1 = klt.pinned_empty((B, rows, cols), np.uint8)
    imgs2 = klt.pinned_empty((B, rows, cols), np.uint8)
    for b in range(B):
        src = base[b % len(base)]
        y0, x0 = int(rng.integers(0, src.shape[0] - rows + 1)), int(rng.integers(0, src.shape[1] - cols + 1))
        imgs1[b] = src[y0:y0 + rows, x0:x0 + cols]
        imgs2[b] = np.roll(imgs1[b], 3, axis=1)
    imgs1[5] = 77                                   # a flat image: no corners at all
    kp1 = klt.pinned_empty((B, nk, 2), np.float32)
    kp1[:] = np.stack([rng.uniform(-5, cols + 5, (B, nk)), rng.uniform(-5, rows + 5, (B, nk))], axis=2)
    counts = rng.integers(0, nk + 1, B).astype(np.int32)
    batch = tracker.batch(B, rows, cols, nk, levels=3)
    batch.set_feature_counts(counts)
    batch.upload(imgs1, imgs2, kp1, kp1)
    for image_set, imgs in ((0, imgs1), (1, imgs2)):
        pts, cnt, sc = batch.detect_features(image_set, n, 0.01, md, exclude_keypoints=exclude, exclude_half=10.0)
        for b in range(B):
            ex = kp1[b, :counts[b]] if exclude and counts[b] else None
            ref, rsc = tracker.detect_features(np.ascontiguousarray(imgs[b]), n, 0.01, md, exclude=ex)
            assert cnt[b] == ref.shape[0], (image_set, b)
            assert np.array_equal(pts[b, :cnt[b]], ref) and np.array_equal(sc[b, :cnt[b]].view(np.uint32), rsc.view(np.uint32))
            assert not pts[b, cnt[b]:].any()
        assert image_set == 1 or cnt[5] == 0
    batch.close()


@pytest.mark.gpu
def test_gpu_batched_detection_in_several_workspace_passes_and_on_1080p(tracker, monkeypatch):
    """More images than one pass through the workspace holds (LEGO_KLT_DETECT_CHUNK = 5 of 13), and a 1920x1080 image
    (strips and row chunks that do not divide the image) against the numpy oracle."""
    import lego_slam_b200 as klt
    B, rows, cols = 13, 97, 131
    rng = np.random.default_rng(4)
    imgs = klt.pinned_empty((B, rows, cols), np.uint8)
    src = _images()[2]
    for b in range(B):
        imgs[b] = np.roll(src, (int(rng.integers(0, rows)), int(rng.integers(0, cols))), axis=(0, 1))
    kp = klt.pinned_empty((B, 4, 2), np.float32)
    kp[:] = 5
    batch = tracker.batch(B, rows, cols, 4, levels=2)
    batch.upload(imgs, imgs, kp, kp)
    monkeypatch.setenv("LEGO_KLT_DETECT_CHUNK", "5")
    pts, cnt, sc = batch.detect_features(1, 40, 0.01, 7.0)
    monkeypatch.delenv("LEGO_KLT_DETECT_CHUNK")
    for b in range(B):
        ref, rsc = g.good_features_to_track(np.ascontiguousarray(imgs[b]), 40, 0.01, 7.0)
        assert cnt[b] == ref.shape[0] and np.array_equal(pts[b, :cnt[b]], ref)
        assert np.array_equal(sc[b, :cnt[b]].view(np.uint32), rsc.view(np.uint32))
    batch.close()
    big, _, _ = synth.stereo_pair(1080, 1920, 3)
    pts, sc = tracker.detect_features(big, 500, 0.01, 15.0)
    e_gpu = tracker.debug_read_eig(1080, 1920)
    e, _ = g.corner_min_eigen_val(big)
    assert np.array_equal(e_gpu.view(np.uint32), e.view(np.uint32))
    ref, rsc = g.good_features_to_track(big, 500, 0.01, 15.0)
    assert np.array_equal(pts, ref) and np.array_equal(sc.view(np.uint32), rsc.view(np.uint32))
    one = tracker.batch(1, 1080, 1920, 4, levels=2)
    k1 = klt.pinned_empty((1, 4, 2), np.float32)
    k1[:] = 9
    one.upload(big[None], big[None], k1, k1)
    bp, bc, _ = one.detect_features(0, 500, 0.01, 15.0)
    assert bc[0] == ref.shape[0] and np.array_equal(bp[0, :bc[0]], ref)
    one.close()


@pytest.mark.gpu
def test_gpu_detect_track_triangulate_chain_stays_on_the_device(tracker, oracle):
    """lego_klt_batch_detect_features -> lego_klt_batch_use_detected_features -> lego_klt_batch_run: the detected
    corners become the source keypoints where they lie in HBM.  Same bytes as fetching the corners, uploading them as
    keypoints of a ragged batch and running; one pair also against the CPU oracle."""
    import lego_slam_b200 as klt
    B, rows, cols, n = 6, 188, 620, 200
    imgs1 = klt.pinned_empty((B, rows, cols), np.uint8)
    imgs2 = klt.pinned_empty((B, rows, cols), np.uint8)
    for b in range(B):
        imgs1[b], imgs2[b], _ = synth.stereo_pair(rows, cols, 40 + b)
    zeros = klt.pinned_empty((B, n, 2), np.float32)
    zeros[:] = 0
    batch = tracker.batch(B, rows, cols, n, levels=4)
    batch.upload(imgs1, imgs2, zeros, zeros)
    pts, cnt, _ = batch.detect_features(0, n, 0.01, 12.0)
    assert cnt.min() > 20
    batch.use_detected_features()
    batch.run(klt.make_params(kernel=klt.KERNEL_LANE))
    o1, s1, st1 = batch.download()
    assert int(st1.n_features) == int(cnt.sum())
    # the host round trip
    other = tracker.batch(B, rows, cols, n, levels=4)
    other.set_feature_counts(cnt)
    kp = klt.pinned_empty((B, n, 2), np.float32)
    kp[:] = pts
    other.upload(imgs1, imgs2, kp, kp)
    other.run(klt.make_params(kernel=klt.KERNEL_LANE))
    o2, s2, st2 = other.download()
    used = np.arange(n)[None, :] < cnt[:, None]
    assert np.array_equal(o1.view(np.uint32)[used], o2.view(np.uint32)[used]) and np.array_equal(s1[used], s2[used])
    assert [int(v) for v in st1.gn_iters][:4] == [int(v) for v in st2.gn_iters][:4]
    ref, rs, _ = oracle.track(imgs1[2], imgs2[2], pts[2, :cnt[2]], pts[2, :cnt[2]])
    from parity_util import assert_parity
    assert_parity(o1[2, :cnt[2]], s1[2, :cnt[2]], ref, rs, cols, rows, "chain pair 2")
    batch.close()
    other.close()


@pytest.mark.gpu
def test_detected_features_track(tracker, oracle):
    """The frontend's sequence on the device: detect on the left image, track into the right one."""
    import lego_slam_b200 as klt
    L, R, _ = synth.stereo_pair(376, 1241, 4)
    kp, _ = tracker.detect_features(L, 150, 0.01, 20.0)
    out, ok, _ = tracker.track(L, R, kp, kp, klt.make_params())
    ref, rok, _ = oracle.track(L, R, kp, kp)
    assert np.array_equal(ok, rok) and np.abs(out - ref).max() <= 1e-3 and ok.mean() > 0.9
