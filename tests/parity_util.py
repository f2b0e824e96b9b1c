"""Parity report GPU-vs-oracle (BASELINE.json north_star): flags bit-identical except borderline
features (final position within 1e-4 px of an image border -- the reference has no error threshold,
the in-image test of src/algorithm.cpp:123 is its only data-dependent flag decision besides NaN);
positions within 1e-3 px."""
import numpy as np

POS_TOL = 1e-3     # px, stated in BASELINE.json
BORDER_TOL = 1e-4  # px


def border_distance(kp, cols, rows):
    x, y = kp[..., 0].astype(np.float64), kp[..., 1].astype(np.float64)
    return np.minimum.reduce([np.abs(x), np.abs(y), np.abs(x - cols), np.abs(y - rows)])


def parity_report(gpu_kp, gpu_succ, ref_kp, ref_succ, cols, rows):
    gpu_kp, ref_kp = np.asarray(gpu_kp).reshape(-1, 2), np.asarray(ref_kp).reshape(-1, 2)
    gpu_succ, ref_succ = np.asarray(gpu_succ).reshape(-1), np.asarray(ref_succ).reshape(-1)
    d = np.abs(gpu_kp.astype(np.float64) - ref_kp.astype(np.float64)).max(axis=1) if len(gpu_kp) else np.zeros(0)
    flag_diff = np.nonzero(gpu_succ.astype(bool) != ref_succ.astype(bool))[0]
    borderline = [int(i) for i in flag_diff if border_distance(ref_kp[i], cols, rows) < BORDER_TOL]
    hard_flag = [int(i) for i in flag_diff if int(i) not in borderline]
    return {
        "n": int(len(d)),
        "bit_identical": int((gpu_kp.view(np.uint32) == ref_kp.view(np.uint32)).all(axis=1).sum()) if len(d) else 0,
        "max_abs_diff_px": float(d.max()) if len(d) else 0.0,
        "n_over_tol": int((d > POS_TOL).sum()),
        "flag_mismatch_borderline": borderline,
        "flag_mismatch_hard": hard_flag,
    }


def assert_parity(gpu_kp, gpu_succ, ref_kp, ref_succ, cols, rows, what=""):
    rep = parity_report(gpu_kp, gpu_succ, ref_kp, ref_succ, cols, rows)
    assert not rep["flag_mismatch_hard"], f"{what}: success flags differ: {rep}"
    assert rep["n_over_tol"] == 0, f"{what}: positions differ by more than {POS_TOL} px: {rep}"
    return rep
