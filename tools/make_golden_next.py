"""Generates tests/golden/next_rows.npz (+ entries in next_rows.json): frozen vectors for the rows either side of the
KLT path (SURVEY.md 8f N2 frame ingest, N3 triangulation).  Run in the BUILD container (cv2 importable).

Pins: cv2.resize(..., 0.5, 0.5, INTER_NEAREST) itself for the ingest (the call at src/dataset.cpp:75-77); the
reference's own unit test (test/legoslam_test_triangulation.cpp:5-23) plus LAPACK outputs, frozen, for triangulation.

    python tools/make_golden_next.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ingest_np, triangulation_np as tri  # noqa: E402
import test_triangulation as T  # noqa: E402  (the synthetic rig and the reference's KAT live with the tests)

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import cv2
    out, meta = {}, {"cv2_version": cv2.__version__, "ingest": {}, "triangulation": {}}
    for name, (rows, cols, seed) in {"kitti_full_1241x376": (376, 1241, 21), "odd_2483x751": (751, 2483, 22),
                                     "tiny_5x3": (3, 5, 23)}.items():
        img = np.random.default_rng(seed).integers(0, 256, size=(rows, cols), dtype=np.uint8)
        ref = cv2.resize(img, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_NEAREST)
        assert np.array_equal(ref, ingest_np.downscale_half_nearest(img)), name
        meta["ingest"][name] = {"rows": rows, "cols": cols, "seed": seed, "out_shape": list(ref.shape), "sha256": sha(ref)}
    # the reference's unit test
    pw, poses, points = T.reference_kat()
    est, ok = tri.triangulation(poses, points)
    assert ok and np.abs(est - pw).max() < 0.01
    out.update(kat_poses=poses, kat_points=points[:, :2], kat_pt=est, kat_ok=np.uint8(ok))
    # noisy stereo tracks on the KITTI rig: pixels in, world points + verdicts out
    n = 4096
    _, kl, kr, left, right = T.synthetic_tracks(n, seed=31)
    k = T.KITTI
    pts = np.stack([tri.pixel2camera(kl, k["fx"], k["fy"], k["cx"], k["cy"]),
                    tri.pixel2camera(kr, k["fx"], k["fy"], k["cx"], k["cy"])], axis=1)
    pt, okv, ratio = tri.triangulation_batch([left, right], pts, 1e-3)
    out.update(stereo_kl=kl, stereo_kr=kr, stereo_left=left, stereo_right=right, stereo_pt=pt, stereo_ok=okv,
               stereo_ratio=ratio)
    meta["triangulation"] = {"n": n, "n_ok": int(okv.sum()), "thr": 1e-3, "kitti": k}
    np.savez_compressed(os.path.join(OUT, "next_rows.npz"), **out)
    with open(os.path.join(OUT, "next_rows.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
