"""One 2000-feature call on the warp kernel (the reference's call pattern), for profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lego_slam_b200 as klt
from lego_slam_b200 import synth
trk = klt.Tracker(0)
L, R, kp1, kp2, _ = synth.stereo_case(376, 1241, 2000, seed=2)
a = trk.image(376, 1241).upload(L)
b = trk.image(376, 1241).upload(R)
for _ in range(4):
    out, succ, st = trk.track_images(a, b, kp1, kp2, klt.make_params(kernel=klt.KERNEL_WARP))
print(int(succ.sum()), list(st.gn_iters)[:4], st.n_slow_path)
