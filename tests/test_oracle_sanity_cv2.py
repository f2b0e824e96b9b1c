"""Sanity (NOT parity): the oracle's forward-mode tracks against OpenCV's own pyramidal LK
(cv2.calcOpticalFlowPyrLK, the alternative path of the reference's frontend, src/frontend_g2o.cpp:370-450) and against
the analytic ground truth of the synthetic stereo pair.  Different algorithms (window, levels, iteration rule), so the
bar is statistical: both land on the same match to a fraction of a pixel on easy inputs."""
import numpy as np
import pytest

from lego_slam_b200 import synth
from oracle import binding as ob

cv2 = pytest.importorskip("cv2")


def test_oracle_tracks_agree_with_opencv_pyr_lk_and_ground_truth():
    L, R, kp1, kp2, truth = synth.stereo_case(376, 1241, 400, seed=21, min_dist=12, guess="noisy")
    out, ok, _ = ob.track(L, R, kp1, kp2, ob.make_params())
    p1 = kp1.reshape(-1, 1, 2).astype(np.float32)
    p2 = kp2.reshape(-1, 1, 2).astype(np.float32).copy()
    cv_out, cv_ok, _ = cv2.calcOpticalFlowPyrLK(
        L, R, p1, p2, winSize=(11, 11), maxLevel=3,
        criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01), flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    cv_out, cv_ok = cv_out.reshape(-1, 2), cv_ok.reshape(-1).astype(bool)
    both = ok.astype(bool) & cv_ok
    assert both.mean() > 0.9
    err_oracle = np.linalg.norm(out[both] - truth[both], axis=1)
    err_cv = np.linalg.norm(cv_out[both] - truth[both], axis=1)
    diff = np.linalg.norm(out[both] - cv_out[both], axis=1)
    # the 7x7 Gauss-Newton tracker and OpenCV's 11x11 one find the same disparity
    assert np.median(err_oracle) < 0.25 and np.median(err_cv) < 0.25
    assert np.median(diff) < 0.25
    assert (diff < 1.0).mean() > 0.85


def test_inverse_mode_of_the_reference_is_not_a_tracker():
    """Guards the bug-compatibility of the inverse mode: with its stale Jacobian it does NOT converge to
    the match the forward mode finds -- anyone 'fixing' it would make this test fail together with the parity tests."""
    L, R, kp1, kp2, truth = synth.stereo_case(376, 1241, 300, seed=22, min_dist=12)
    fwd, _, _ = ob.track(L, R, kp1, kp2, ob.make_params())
    inv, _, _ = ob.track(L, R, kp1, kp2, ob.make_params(inverse=True))
    e_f = np.median(np.linalg.norm(fwd - truth, axis=1))
    e_i = np.median(np.linalg.norm(inv - truth, axis=1))
    assert e_f < 0.5 < e_i
