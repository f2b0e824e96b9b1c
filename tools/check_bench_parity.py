"""Parity of the fast kernels against the on-GPU EXACT kernel (bit-identical to the oracle) on the full
bench workload: flags must be equal, positions within 1e-3 px."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lego_slam_b200 as klt

B, n = int(os.environ.get("PAIRS", 256)), 2000
base = bench.make_workload(B, n, 64, 1000)
imgs1, imgs2, kp1, kp2 = bench.fill_batch(base, B, n, klt.pinned_empty)
trk = klt.Tracker(0)
res = {}
for name, k in (("exact", klt.KERNEL_EXACT), ("lane", klt.KERNEL_AUTO), ("warp", klt.KERNEL_WARP)):
    batch = trk.batch(B, bench.ROWS, bench.COLS, n, levels=4)   # separate output buffers per kernel
    batch.upload(imgs1, imgs2, kp1, kp2)
    batch.run(klt.make_params(kernel=k))
    o, s, st = batch.download()
    res[name] = (o.copy(), s.copy(), [int(v) for v in st.gn_iters][:4], int(st.n_success))
    print(name, "iters", res[name][2], "succ", res[name][3], "slow", st.n_slow_path, "deferred", st.n_deferred)
eo, es = res["exact"][:2]
for name in ("lane", "warp"):
    o, s = res[name][:2]
    d = np.abs(o.astype(np.float64) - eo).max(axis=2)
    bad = np.argwhere((d > 1e-3) | (s != es))
    print(name, "max diff", d.max(), "bit-identical", float((o.view(np.uint32) == eo.view(np.uint32)).all(axis=2).mean()), "bad", len(bad))
    for b, i in bad[:8]:
        print("   pair", b, "feat", i, "kp1", kp1[b, i], "exact", eo[b, i], es[b, i], name, o[b, i], s[b, i])
