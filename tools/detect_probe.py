"""Feature-detection call for profiling: python tools/detect_probe.py [N_CORNERS] [MIN_DIST]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lego_slam_b200 as klt
from lego_slam_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
md = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
L, R, _ = synth.stereo_pair(376, 1241, 1)
trk = klt.Tracker(0)
h = trk.image(376, 1241, 4).upload(L)
for _ in range(3):
    pts, _ = trk.detect_features(h, n, 0.01, md)
t0 = time.perf_counter()
for _ in range(20):
    pts, _ = trk.detect_features(h, n, 0.01, md)
print(f"{n} corners requested, {len(pts)} found, min distance {md}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per call")
