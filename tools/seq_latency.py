"""Latency of one track call (image handles, solver only) by kernel and feature count."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lego_slam_b200 as klt
from lego_slam_b200 import synth

trk = klt.Tracker(0)
counts = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [150, 500, 2000, 5000, 20000]
for n in counts:
    L, R, kp1, kp2, _ = synth.stereo_case(376, 1241, n, seed=2, min_dist=3 if n > 5000 else None)
    a = trk.image(376, 1241).upload(L)
    b = trk.image(376, 1241).upload(R)
    row = [f"n={kp1.shape[0]:6d}"]
    kernels = (("lane", klt.KERNEL_LANE), ("warp", klt.KERNEL_WARP), ("patch", klt.KERNEL_PATCH))
    if len(sys.argv) > 2:
        kernels = tuple(kv for kv in kernels if kv[0] in sys.argv[2].split(","))
    for name, k in kernels:
        p = klt.make_params(kernel=k)
        for _ in range(5):
            trk.track_images(a, b, kp1, kp2, p)
        t0 = time.perf_counter()
        for _ in range(50):
            trk.track_images(a, b, kp1, kp2, p)
        row.append(f"{name} {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
        t0 = time.perf_counter()
        for _ in range(50):
            trk.track_images(a, b, kp1, kp2, p, want_stats=False)
        row.append(f"(no stats {(time.perf_counter() - t0) / 50 * 1e3:.3f})")
    print("  ".join(row))
