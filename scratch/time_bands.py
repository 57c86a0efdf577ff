"""sid_run_pair (pinned images) by number of upload bands, default kernel: python scratch/time_bands.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sea_ice_drift_b200 import _lib, synthetic as syn
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=0)
s = cfg["img_size"]; angles = cfg["angles"]
img1p = torch.from_numpy(img1).pin_memory().numpy(); img2p = torch.from_numpy(img2).pin_memory().numpy()
ctx = _lib.Context(0)
def timeit(label, fn, reps=8):
    fn(); fn(); ctx.synchronize(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    print("%-46s best %7.3f ms  median %7.3f ms" % (label, min(ts) * 1e3, float(np.median(ts)) * 1e3), flush=True)
timeit("set_pair (pinned, 2 x 108 MB) + sync", lambda: (ctx.set_pair(img1p, img2p), ctx.synchronize()))
timeit("run (resident pair, host points / results)", lambda: ctx.run(c1, r1, c2, r2, b, s, angles, 0.0))
timeit("run_pair, default band plan, pinned", lambda: ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0))
for nb in (6, 8, 10, 12):
    os.environ["SID_BANDS"] = str(nb)
    timeit("run_pair, %d equal band(s), pinned" % nb, lambda: ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0))
del os.environ["SID_BANDS"]
for plan in ("6,6,6,6,6,5,4,3,2,1", "8,8,8,8,6,4,3,2,1", "4,6,6,6,6,6,4,3,2,1,0.5", "6,6,6,6,6,6,4,2,1", "8,8,8,8,8,4,2,1", "10,10,10,8,4,2,1",
             "5,5,5,5,5,5,5,4,3,2,1", "3,5,6,6,6,6,5,4,3,2,1,0.5"):
    os.environ["SID_BAND_PLAN"] = plan
    timeit("run_pair, plan %s" % plan, lambda: ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0))
