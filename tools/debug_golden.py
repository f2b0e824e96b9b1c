import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lego_slam_b200 as klt
g = dict(np.load("tests/golden/solver_620x188.npz"))
meta = json.load(open("tests/golden/golden.json"))["solver"]
trk = klt.Tracker(0)
for name in ("fwd", "fwd_noinit", "fwd_1layer"):
    kw = dict(meta[name]); iters = kw.pop("gn_iters"); kw.pop("n_success")
    for kernel in (klt.KERNEL_WARP, klt.KERNEL_LANE):
        out, succ, st = trk.track(g["left"], g["right"], g["kp1"], g["kp2"], klt.make_params(kernel=kernel, **kw))
        d = np.abs(out.astype(np.float64) - g[name + "_kp2"]).max(axis=1)
        bad = np.nonzero((d > 1e-3) | (succ != g[name + "_succ"]))[0]
        print(name, kernel, "bad", bad.tolist(), "iters", list(st.gn_iters)[:4], iters, "slow", st.n_slow_path, "deferred", st.n_deferred, list(st.defer_reason))
        for i in bad[:6]:
            print("   ", i, g["kp1"][i], g["kp2"][i], "got", out[i], succ[i], "want", g[name + "_kp2"][i], g[name + "_succ"][i])
