"""Device-resident solver throughput by patch size and kernel on the bench workload (64 pairs by default)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lego_slam_b200 as klt

B, n = int(os.environ.get("PAIRS", 128)), 2000
base = bench.make_workload(B, n, 32, 1000)
imgs1, imgs2, kp1, kp2 = bench.fill_batch(base, B, n, klt.pinned_empty)
trk = klt.Tracker(0)
batch = trk.batch(B, bench.ROWS, bench.COLS, n, levels=4)
batch.upload(imgs1, imgs2, kp1, kp2)
for lo, hi in ((-3, 3), (-4, 3), (-5, 5)):
    for name, k in (("lane", klt.KERNEL_LANE), ("warp", klt.KERNEL_WARP)):
        p = klt.make_params(patch_lo=lo, patch_hi=hi, kernel=k)
        batch.run(p); trk.sync()
        t0 = time.perf_counter()
        reps = 5 if name == "lane" else 2
        for _ in range(reps):
            batch.run(p)
        trk.sync()
        ms = (time.perf_counter() - t0) / reps * 1e3
        print(f"patch {lo}..{hi} {name}: {ms:.3f} ms per {B * n} features -> {B * n / ms * 1e3:.4g} tracks/s")
