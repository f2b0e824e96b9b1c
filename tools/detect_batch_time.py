import sys, time, numpy as np
sys.path.insert(0,'.')
import bench, lego_slam_b200 as klt
B=256
base=bench.make_workload(150,16,1000)
imgs1,imgs2,kp1,kp2=bench.fill_batch(base,B,150,klt.pinned_empty)
trk=klt.Tracker(0); fb=trk.batch(B,bench.ROWS,bench.COLS,150,levels=4)
fb.upload(imgs1,imgs2,kp1,kp2)
for n,md in ((150,20.0),(2000,5.0)):
    fb2 = fb if n<=150 else None
    for _ in range(2): fb.detect_features(0,n,0.01,md)
    t0=time.perf_counter()
    for _ in range(5): pts,cnt,_=fb.detect_features(0,n,0.01,md)
    print(n,md,"ms per 256 images",(time.perf_counter()-t0)/5*1e3, "corners", cnt.mean())
