"""finer band plans and staging thread counts: python scratch/time_bands2.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sea_ice_drift_b200 import _lib, synthetic as syn
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=0)
s = cfg["img_size"]; angles = cfg["angles"]
img1p = torch.from_numpy(img1).pin_memory().numpy(); img2p = torch.from_numpy(img2).pin_memory().numpy()
ctx = _lib.Context(0)
def timeit(label, fn, reps=10):
    fn(); fn(); ctx.synchronize(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    print("%-58s best %7.3f ms  median %7.3f ms" % (label, min(ts) * 1e3, float(np.median(ts)) * 1e3), flush=True)
timeit("run_pair, default band plan, pinned", lambda: ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0))
for plan in ("6,6,6,6,6,5,4,3,2,1", "6,6,6,6,6,5,4,3,2,1,0.5", "6,6,6,6,6,5,4,3,2,1,0.5,0.25", "7,7,7,7,6,5,4,3,2,1,0.5", "8,8,8,7,6,5,4,3,2,1,0.5",
             "6,6,6,6,6,5,4,3,2,1.5,1,0.5", "4,6,6,6,6,5,4,3,2,1", "6,6,6,6,6,6,5,4,3,2,1,0.5", "10,9,8,7,6,5,4,3,2,1,0.5", "12,10,8,6,5,4,3,2,1,0.5"):
    os.environ["SID_BAND_PLAN"] = plan
    timeit("run_pair, plan %s" % plan, lambda: ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0))
del os.environ["SID_BAND_PLAN"]
for nt in (4, 8, 12, 16, 24):
    os.environ["SID_UPLOAD_THREADS"] = str(nt)
    timeit("run_pair, pageable images, %d staging threads" % nt, lambda: ctx.run_pair(img1, img2, c1, r1, c2, r2, b, s, angles, 0.0), reps=6)
del os.environ["SID_UPLOAD_THREADS"]
for plan in ("1,1,1,1,1,1,1,1", "6,6,6,6,6,5,4,3,2,1", "3,4,5,6,6,6,5,4,3,2,1", "2,3,4,5,6,6,6,5,4,3,2,1"):
    os.environ["SID_BAND_PLAN"] = plan
    timeit("run_pair, pageable, plan %s" % plan, lambda: ctx.run_pair(img1, img2, c1, r1, c2, r2, b, s, angles, 0.0), reps=6)
import multiprocessing; print("host cores", multiprocessing.cpu_count())
