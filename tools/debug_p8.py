import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lego_slam_b200 as klt
from lego_slam_b200 import synth
trk = klt.Tracker(0)
L, R, kp1, kp2, _ = synth.stereo_case(188, 620, 158, seed=1, min_dist=10)
for levels in (1, 2, 4):
    pe = klt.make_params(levels=levels, patch_lo=-4, patch_hi=3, kernel=klt.KERNEL_EXACT)
    pl = klt.make_params(levels=levels, patch_lo=-4, patch_hi=3, kernel=klt.KERNEL_LANE)
    eo, es, est = trk.track(L, R, kp1, kp2, pe)
    lo, ls, lst = trk.track(L, R, kp1, kp2, pl)
    d = np.abs(eo - lo).max(axis=1)
    bad = np.where(d > 1e-3)[0]
    print("levels", levels, "bad", len(bad), "iters", list(est.gn_iters)[:levels], list(lst.gn_iters)[:levels], "slow", lst.n_slow_path)
    for i in bad[:10]:
        print("   ", i, kp1[i], "x&15", int(kp1[i][0]) & 15, "x&7", int(kp1[i][0]) & 7, "exact", eo[i], "lane", lo[i])
