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
for nb in (4, 6, 8, 10, 12, 16):
    os.environ["SID_BANDS"] = str(nb)
    timeit("run_pair, %d band(s), pinned" % nb, lambda: ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0))
