"""Generates tests/golden/*.npz|json (run in the BUILD container, where cv2 is importable).

The reference ships no golden vectors for the KLT path, so they are generated here from the reference ITSELF:
(1) OpenCV for the pyramid (cv2.resize, the library the reference calls at src/algorithm.cpp:147-150);
(2) the solver vectors are the outputs of oracle/_ref -- the reference's own translation unit
    (/root/reference/src/algorithm.cpp + include/legoslam/algorithm.h compiled unmodified on stand-in headers,
    oracle/build_ref.py) -- for every variant the reference text can express ("source": "reference_tu"; with the
    patch / level literals substituted: "reference_tu_parametrised"); the asymmetric 8x8 patch is not expressible
    in the reference's loops and comes from the oracle ("oracle").  At generation time the oracle and the
    independent numpy restatement are asserted bit-equal to those outputs; iteration counts come from the oracle
    (the reference exports none).

    python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lego_slam_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402
from oracle import klt_oracle_np as onp  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pyramid_hashes():
    import cv2
    cases = {}
    for name, (rows, cols, levels, seed) in {
        "kitti_1241x376_L4": (376, 1241, 4, 11), "kitti_half_620x188_L4": (188, 620, 4, 12),
        "hd_1920x1080_L5": (1080, 1920, 5, 13), "odd_333x217_L3": (217, 333, 3, 14),
        "tiny_37x23_L3": (23, 37, 3, 15),
    }.items():
        img = np.random.default_rng(seed).integers(0, 256, size=(rows, cols), dtype=np.uint8)
        cur, hs = img, []
        pyr = ob.build_pyramid(img, levels)
        for l in range(1, levels):
            cur = cv2.resize(cur, (int(cur.shape[1] * 0.5), int(cur.shape[0] * 0.5)))
            assert np.array_equal(cur, pyr[l]), (name, l)
            hs.append({"shape": list(cur.shape), "sha256": sha(cur)})
        cases[name] = {"rows": rows, "cols": cols, "levels": levels, "seed": seed, "levels_sha": hs,
                       "cv2_version": cv2.__version__}
    return cases


def solver_vectors():
    rows, cols, n = 188, 620, 150
    left, right, kp1, kp2, truth = synth.stereo_case(rows, cols, n, seed=1, min_dist=10)
    # put a few awkward features in: near borders, outside, on a binade boundary
    extra = np.array([[3.0, 3.0], [cols - 2.5, rows - 2.5], [cols - 0.5, 40.0], [100.0, rows - 0.25],
                      [-2.0, 50.0], [127.999, 63.999], [255.5, 31.75], [511.9996, 100.0]], np.float32)
    kp1 = np.concatenate([kp1, extra]).astype(np.float32)
    kp2 = np.concatenate([kp2, extra + np.float32(0.3)]).astype(np.float32)
    out = {"left": left, "right": right, "kp1": kp1, "kp2": kp2}
    variants = {
        "fwd": dict(levels=4), "inv": dict(levels=4, inverse=True), "fwd_noinit": dict(levels=4, has_initial=False),
        "fwd_1layer": dict(levels=1), "fwd_8x8": dict(levels=4, patch_lo=-4, patch_hi=3),
        "fwd_11x11": dict(levels=3, patch_lo=-5, patch_hi=5), "inv_11x11": dict(levels=3, patch_lo=-5, patch_hi=5, inverse=True),
    }
    meta = {}
    for name, kw in variants.items():
        p = ob.make_params(**kw)
        o, s, st = ob.track(left, right, kp1, kp2, p)
        source = "oracle"
        if p.patch_lo == -p.patch_hi:
            hp, lv = p.patch_hi, (4 if p.levels == 1 else p.levels)
            ro, rs = rb.track(left, right, kp1, kp2, inverse=bool(p.inverse), has_initial=bool(p.has_initial),
                              layers=p.levels, half_patch=hp, pyramids=lv)
            assert np.array_equal(ro.view(np.uint32), o.view(np.uint32)) and np.array_equal(rs, s), name
            o, s = ro, rs
            source = "reference_tu" if (hp, lv) == (3, 4) else "reference_tu_parametrised"
        out[f"{name}_kp2"] = o
        out[f"{name}_succ"] = s
        meta[name] = dict(kw, gn_iters=[int(v) for v in st.gn_iters][:p.levels], n_success=int(st.n_success),
                          source=source)
        # cross-check the first 40 + the awkward ones against the numpy restatement
        sel = np.r_[0:40, n:n + extra.shape[0]]
        p1 = ob.build_pyramid(left, p.levels)
        p2 = ob.build_pyramid(right, p.levels)
        a, sa, _ = onp.track(p1, p2, kp1[sel], kp2[sel], levels=p.levels, lo=p.patch_lo, hi=p.patch_hi,
                             inverse=bool(p.inverse), has_initial=bool(p.has_initial))
        assert np.array_equal(a.view(np.uint32), o[sel].view(np.uint32)), name
        assert np.array_equal(sa, s[sel]), name
    np.savez_compressed(os.path.join(OUT, "solver_620x188.npz"), **out)
    return meta


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    meta = {"pyramid": pyramid_hashes(), "solver": solver_vectors()}
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta["solver"], indent=1))
