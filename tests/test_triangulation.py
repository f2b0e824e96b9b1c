"""Triangulation of tracked features: legoslam::triangulation, include/legoslam/algorithm.h:11-34.

The reference pins this function with its ONLY unit test, test/legoslam_test_triangulation.cpp:5-23 -- restated here
as the known-answer test of both the numpy oracle (CPU) and the sm_100a kernel (through the C ABI).
Tolerances (fp64 SVD by different but backward-stable algorithms): world points relative 1e-9; the bool is
identical except where S[3]/S[2] lies within 1e-9 (relative) of the threshold.
"""
import numpy as np
import pytest

import lego_slam_b200 as klt
from oracle import triangulation_np as tri

KITTI = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, baseline=0.537)   # KITTI-00 P0/P1 (calib.txt)


def reference_kat():
    """test/legoslam_test_triangulation.cpp:5-23.  Eigen::Quaterniond(0, 0, 0, 1) is w=0, z=1: 180 deg about z."""
    pt_world = np.array([30.0, 20.0, 10.0])
    rot = tri.quat_to_rot(0, 0, 0, 1)
    poses = np.stack([tri.se3_matrix3x4(rot, t) for t in ([0, 0, 0], [0, -10, 0], [0, 10, 0])])
    points = []
    for m in poses:
        pc = m[:, :3] @ pt_world + m[:, 3]
        points.append(pc / pc[2])
    return pt_world, poses, np.array(points)


def stereo_rig():
    """Frontend's poses = {camera_left->pose(), camera_right->pose()} for a rectified rig (src/dataset.cpp:22-48:
    t = K^-1 * P[:, 3], pose = SE3(I, t))."""
    k = KITTI
    left = tri.se3_matrix3x4(np.eye(3), [0, 0, 0])
    right = tri.se3_matrix3x4(np.eye(3), [-k["baseline"], 0, 0])
    return left, right


def synthetic_tracks(n, seed, noise_px=0.3):
    rng = np.random.default_rng(seed)
    k = KITTI
    left, right = stereo_rig()
    P = np.stack([rng.uniform(-20, 20, n), rng.uniform(-3, 2, n), rng.uniform(4, 80, n)], axis=1)
    def project(m):
        pc = P @ m[:, :3].T + m[:, 3]
        return np.stack([k["fx"] * pc[:, 0] / pc[:, 2] + k["cx"], k["fy"] * pc[:, 1] / pc[:, 2] + k["cy"]], axis=1)
    kl = (project(left) + rng.normal(0, noise_px, (n, 2))).astype(np.float32)
    kr = (project(right) + rng.normal(0, noise_px, (n, 2))).astype(np.float32)
    return P, kl, kr, left, right


def assert_tri_parity(pt, ok, ref_pt, ref_ok, ratio, thr, what=""):
    borderline = np.abs(ratio - thr) <= 1e-9 * thr
    assert np.array_equal(ok[~borderline], ref_ok[~borderline]), f"{what}: verdicts differ"
    fin = np.all(np.isfinite(ref_pt), axis=1)
    err = np.abs(pt[fin] - ref_pt[fin]) / np.maximum(np.abs(ref_pt[fin]), 1.0)
    assert err.size == 0 or err.max() <= 1e-9, f"{what}: max relative deviation {err.max():.3e}"


# ------------------------------------------------------------------ CPU: the oracle against the reference's KAT
def test_oracle_reproduces_the_reference_unit_test():
    pt_world, poses, points = reference_kat()
    est, ok = tri.triangulation(poses, points)
    assert ok                                                   # EXPECT_TRUE(triangulation(...))
    assert np.all(np.abs(est - pt_world) < 0.01)                # EXPECT_NEAR(..., 0.01) x3
    assert np.all(np.abs(est - pt_world) < 1e-9)


def test_oracle_batch_equals_single_and_recovers_noise_free_points():
    P, kl, kr, left, right = synthetic_tracks(64, seed=5)
    k = KITTI
    pts = np.stack([tri.pixel2camera(kl, k["fx"], k["fy"], k["cx"], k["cy"]),
                    tri.pixel2camera(kr, k["fx"], k["fy"], k["cx"], k["cy"])], axis=1)
    bp, bok, _ = tri.triangulation_batch([left, right], pts, 1e-3)
    for i in range(8):
        sp, sok = tri.triangulation([left, right], pts[i], 1e-3)
        assert np.allclose(sp, bp[i], rtol=1e-12, atol=0) and sok == bool(bok[i])
    # noise-free stereo tracks triangulate to the generating points; the ratio test passes
    P0, kl0, kr0, _, _ = synthetic_tracks(64, seed=6, noise_px=0.0)
    pts0 = np.stack([tri.pixel2camera(kl0, k["fx"], k["fy"], k["cx"], k["cy"]),
                     tri.pixel2camera(kr0, k["fx"], k["fy"], k["cx"], k["cy"])], axis=1)
    bp0, bok0, _ = tri.triangulation_batch([left, right], pts0, 1e-3)
    assert bok0.all() and np.abs(bp0 - P0).max() < 0.05        # (float32 pixel coordinates)
    # (identical views make S[3]/S[2] a 0/0 of rounding errors: undefined in the reference too, not tested)


# ------------------------------------------------------------------ GPU: the kernel through the C ABI
@pytest.mark.gpu
def test_gpu_reproduces_the_reference_unit_test(tracker):
    pt_world, poses, points = reference_kat()
    est, ok = tracker.triangulation(poses, points[None, :, :2])
    assert ok[0] == 1
    assert np.all(np.abs(est[0] - pt_world) < 0.01)
    assert np.all(np.abs(est[0] - pt_world) < 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("n_views,n", [(2, 5000), (3, 1000), (8, 257)])
def test_gpu_generic_matches_oracle(tracker, n_views, n):
    rng = np.random.default_rng(n_views)
    P = np.stack([rng.uniform(-20, 20, n), rng.uniform(-3, 2, n), rng.uniform(4, 80, n)], axis=1)
    poses = []
    for v in range(n_views):
        q = np.array([1.0, *rng.normal(0, 0.02, 3)])
        poses.append(tri.se3_matrix3x4(tri.quat_to_rot(*q), rng.uniform(-1, 1, 3)))
    poses = np.stack(poses)
    pts = np.empty((n, n_views, 2))
    for v, m in enumerate(poses):
        pc = P @ m[:, :3].T + m[:, 3]
        pts[:, v] = pc[:, :2] / pc[:, 2:3] + rng.normal(0, 2e-4, (n, 2))
    for thr in (1e-3, 1e-2):
        ref_pt, ref_ok, ratio = tri.triangulation_batch(poses, pts, thr)
        pt, ok = tracker.triangulation(poses, pts, thr)
        assert_tri_parity(pt, ok, ref_pt, ref_ok, ratio, thr, f"{n_views} views thr {thr}")
        assert 0 < ok.sum() <= n


@pytest.mark.gpu
def test_gpu_stereo_from_pixels_matches_oracle_and_skips_lost_tracks(tracker):
    n = 20000
    P, kl, kr, left, right = synthetic_tracks(n, seed=11)
    k = KITTI
    cam_l = klt.make_camera(k["fx"], k["fy"], k["cx"], k["cy"], left)
    cam_r = klt.make_camera(k["fx"], k["fy"], k["cx"], k["cy"], right)
    valid = (np.random.default_rng(1).random(n) > 0.1).astype(np.uint8)
    pts = np.stack([tri.pixel2camera(kl, k["fx"], k["fy"], k["cx"], k["cy"]),
                    tri.pixel2camera(kr, k["fx"], k["fy"], k["cx"], k["cy"])], axis=1)
    ref_pt, ref_ok, ratio = tri.triangulation_batch([left, right], pts, 1e-3)
    pt, ok = tracker.triangulation_stereo(cam_l, cam_r, kl, kr, valid, 1e-3)
    v = valid.astype(bool)
    assert_tri_parity(pt[v], ok[v], ref_pt[v], ref_ok[v], ratio[v], 1e-3, "stereo")
    assert not ok[~v].any() and not pt[~v].any()               # no right feature -> nothing triangulated
    pt2, ok2 = tracker.triangulation_stereo(cam_l, cam_r, kl, kr, None, 1e-3)
    assert_tri_parity(pt2, ok2, ref_pt, ref_ok, ratio, 1e-3, "stereo, no mask")


@pytest.mark.gpu
def test_gpu_batch_triangulate_uses_the_tracked_keypoints_in_place(tracker):
    from lego_slam_b200 import synth
    B, n, rows, cols = 3, 400, 188, 620
    cases = [synth.stereo_case(rows, cols, n, seed=40 + b, min_dist=5) for b in range(B)]
    imgs1 = np.stack([c[0] for c in cases]); imgs2 = np.stack([c[1] for c in cases])
    kp1 = np.stack([c[2] for c in cases]).astype(np.float32); kp2 = np.stack([c[3] for c in cases]).astype(np.float32)
    batch = tracker.batch(B, rows, cols, n, levels=4)
    batch.upload(imgs1, imgs2, kp1, kp2)
    batch.run(klt.make_params())
    out, succ, _ = batch.download()
    k = KITTI
    left, right = stereo_rig()
    cam_l = klt.make_camera(k["fx"], k["fy"], k["cx"], k["cy"], left)
    cam_r = klt.make_camera(k["fx"], k["fy"], k["cx"], k["cy"], right)
    pt, ok = batch.triangulate(cam_l, cam_r, 1e-3)
    pt_h, ok_h = tracker.triangulation_stereo(cam_l, cam_r, kp1.reshape(-1, 2), out.reshape(-1, 2), succ.reshape(-1), 1e-3)
    assert np.array_equal(pt.reshape(-1, 3).view(np.uint64), pt_h.view(np.uint64)) and np.array_equal(ok.reshape(-1), ok_h)


# ------------------------------------------------------------------ committed golden vectors (tools/make_golden_next.py)
def _golden_next():
    import json
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(gdir, "next_rows.json")) as f:
        meta = json.load(f)
    return meta, dict(np.load(os.path.join(gdir, "next_rows.npz")))


def test_oracle_reproduces_the_committed_triangulation_vectors():
    meta, g = _golden_next()
    est, ok = tri.triangulation(g["kat_poses"], np.c_[g["kat_points"], np.ones(3)])
    assert ok == bool(g["kat_ok"]) and np.allclose(est, g["kat_pt"], rtol=1e-12, atol=0)
    k = meta["triangulation"]["kitti"]
    pts = np.stack([tri.pixel2camera(g["stereo_kl"], k["fx"], k["fy"], k["cx"], k["cy"]),
                    tri.pixel2camera(g["stereo_kr"], k["fx"], k["fy"], k["cx"], k["cy"])], axis=1)
    pt, okv, ratio = tri.triangulation_batch([g["stereo_left"], g["stereo_right"]], pts, meta["triangulation"]["thr"])
    assert_tri_parity(pt, okv, g["stereo_pt"], g["stereo_ok"], g["stereo_ratio"], meta["triangulation"]["thr"], "golden")
    assert int(okv.sum()) == meta["triangulation"]["n_ok"]


@pytest.mark.gpu
def test_gpu_reproduces_the_committed_triangulation_vectors(tracker):
    meta, g = _golden_next()
    est, ok = tracker.triangulation(g["kat_poses"], g["kat_points"][None])
    assert ok[0] == g["kat_ok"] and np.abs(est[0] - g["kat_pt"]).max() <= 1e-9
    k = meta["triangulation"]["kitti"]
    cam_l = klt.make_camera(k["fx"], k["fy"], k["cx"], k["cy"], g["stereo_left"])
    cam_r = klt.make_camera(k["fx"], k["fy"], k["cx"], k["cy"], g["stereo_right"])
    pt, okv = tracker.triangulation_stereo(cam_l, cam_r, g["stereo_kl"], g["stereo_kr"], None, meta["triangulation"]["thr"])
    assert_tri_parity(pt, okv, g["stereo_pt"], g["stereo_ok"], g["stereo_ratio"], meta["triangulation"]["thr"], "golden stereo")
