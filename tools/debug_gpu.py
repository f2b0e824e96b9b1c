import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lego_slam_b200 as klt
from lego_slam_b200 import synth
from oracle import binding as ob
which = sys.argv[1] if len(sys.argv) > 1 else "all"
L, R, kp1, kp2, _ = synth.stereo_case(188, 620, 150, seed=1, min_dist=10)
trk = klt.Tracker(0)
ref, rs, rst = ob.track(L, R, kp1, kp2)
for name, kernel in (("exact", klt.KERNEL_EXACT), ("warp", klt.KERNEL_WARP), ("lane", klt.KERNEL_LANE)):
    if which not in ("all", name):
        continue
    try:
        out, succ, st = trk.track(L, R, kp1, kp2, klt.make_params(kernel=kernel))
        d = np.abs(out.astype(np.float64) - ref).max()
        print(name, "ok maxdiff", d, "flags equal", np.array_equal(succ, rs), "iters", list(st.gn_iters)[:4], list(rst.gn_iters)[:4],
              "slow", st.n_slow_path, "deferred", st.n_deferred, list(st.defer_reason), "bitident", int((out.view(np.uint32) == ref.view(np.uint32)).all(1).sum()), flush=True)
    except Exception as e:
        print(name, "FAILED", e, flush=True)
        break
