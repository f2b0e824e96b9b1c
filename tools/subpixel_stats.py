import sys, json
sys.path.insert(0,'.')
import numpy as np, bench, lego_slam_b200 as klt, time
import os
B,n=int(os.environ.get("PAIRS",128)),2000
base=bench.make_workload(n,32,1000)
imgs1,imgs2,kp1,kp2=bench.fill_batch(base,B,n,klt.pinned_empty)
trk=klt.Tracker(0); batch=trk.batch(B,bench.ROWS,bench.COLS,n,levels=4)
for name in ("integer","subpixel"):
    if name=="subpixel":
        kp1=kp1+np.random.default_rng(77).uniform(-0.5,0.5,kp1.shape).astype(np.float32); kp1=np.ascontiguousarray(kp1,np.float32); kp2=kp1.copy()
    batch.upload(imgs1,imgs2,kp1,kp2)
    p=klt.make_params()
    for _ in range(5): batch.run(p)
    mp,ms=batch.timings(3)
    _,_,st=batch.download()
    print(name, "pyr %.3f sol %.3f"%(mp,ms), "iters",list(st.gn_iters)[:4],"slow",st.n_slow_path,"deferred",st.n_deferred,"stats",list(st.defer_reason))
