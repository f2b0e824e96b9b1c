"""End-to-end throughput of ONE process driving several GPUs through lego_klt_multi_track (one host thread per device),
beside the one-process-per-GPU figure bench.py reports.   python tools/multi_e2e.py [N_DEVICES] [PAIRS_PER_DEVICE]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lego_slam_b200 as klt  # noqa: E402

import torch  # noqa: E402

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
per = int(sys.argv[2]) if len(sys.argv) > 2 else 256
B, n = ndev * per, 2000
base = bench.make_workload(n, 32, 1000)
imgs1, imgs2, kp1, kp2 = bench.fill_batch(base, B, n, klt.pinned_empty)
succ = klt.pinned_empty((B, n), np.uint8)
io = klt.pinned_empty((B, n, 2), np.float32)
multi = klt.MultiTracker(list(range(ndev)), B, bench.ROWS, bench.COLS, n, levels=4)
p = klt.make_params()
for _ in range(3):
    np.copyto(io, kp2)
    multi.track(imgs1, imgs2, kp1, io, succ, p)
steps = 10
t0 = time.perf_counter()
for _ in range(steps):
    multi.track(imgs1, imgs2, kp1, io, succ, p)      # (kp2 = kp1 on this workload: the tracked points of one step seed the next)
dt = (time.perf_counter() - t0) / steps
print(json.dumps({"what": "lego_klt_multi_track, one process", "devices": ndev, "pairs": B, "ms_per_step": dt * 1e3,
                  "tracks_per_s": B * n / dt, "h2d_gbs_total": (imgs1.nbytes + imgs2.nbytes + kp1.nbytes + io.nbytes) / dt / 1e9}))
