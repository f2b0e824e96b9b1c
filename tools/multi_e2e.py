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
blocks = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]   # 0 = static blocks, else pairs per pulled block
B, n = ndev * per, 2000
base = bench.make_workload(n, 32, 1000)
imgs1, imgs2, kp1, kp2 = bench.fill_batch(base, B, n, klt.pinned_empty)
succ = klt.pinned_empty((B, n), np.uint8)
steps = 8
ring = [klt.pinned_empty((B, n, 2), np.float32) for _ in range(steps)]   # kp2 is in / out: a fresh guess for every timed step
multi = klt.MultiTracker(list(range(ndev)), B, bench.ROWS, bench.COLS, n, levels=4)
p = klt.make_params()
ref = None
for block in blocks:
    multi.set_schedule(block)
    for _ in range(3):
        np.copyto(ring[0], kp2)
        multi.track(imgs1, imgs2, kp1, ring[0], succ, p)
    for io in ring:
        np.copyto(io, kp2)
    t0 = time.perf_counter()
    for io in ring:
        st = multi.track(imgs1, imgs2, kp1, io, succ, p)
    dt = (time.perf_counter() - t0) / steps
    if ref is None:
        ref = (ring[-1].copy(), succ.copy())
    same = bool(np.array_equal(ref[0].view(np.uint32), ring[-1].view(np.uint32)) and np.array_equal(ref[1], succ))
    print(json.dumps({"what": "lego_klt_multi_track, one process", "devices": ndev, "pairs": B, "block_pairs": block,
                      "ms_per_step": dt * 1e3, "tracks_per_s": B * n / dt,
                      "h2d_gbs_total": (imgs1.nbytes + imgs2.nbytes + kp1.nbytes + ring[0].nbytes) / dt / 1e9,
                      "pairs_per_device": multi.last_distribution(),
                      "slowest_device_ms": {"h2d": round(st.ms_h2d, 3), "kernels": round(st.ms_solver + st.ms_pyramid, 3), "d2h": round(st.ms_d2h, 3)}, "same_bytes_as_first_schedule": same}))
