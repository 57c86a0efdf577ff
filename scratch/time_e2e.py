import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sea_ice_drift_b200 import _lib, synthetic as syn
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=0)
n = len(c1); s = cfg["img_size"]; angles = cfg["angles"]
img1p = torch.from_numpy(img1).pin_memory().numpy(); img2p = torch.from_numpy(img2).pin_memory().numpy()
ctx = _lib.Context(0)
def timeit(label, fn, reps=6):
    fn(); fn(); ctx.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    ctx.synchronize(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("%-46s %8.3f ms" % (label, dt * 1e3), flush=True)
timeit("set_pair (pinned, 2 x 108 MB) + sync", lambda: (ctx.set_pair(img1p, img2p), ctx.synchronize()))
timeit("set_pair (pageable) + sync", lambda: (ctx.set_pair(img1, img2), ctx.synchronize()))
ctx.set_pair(img1p, img2p)
timeit("run (resident pair, host points/results)", lambda: ctx.run(c1, r1, c2, r2, b, s, angles, 0.0))
timeit("set_pair + run", lambda: (ctx.set_pair(img1p, img2p), ctx.run(c1, r1, c2, r2, b, s, angles, 0.0)))
for nb in (1, 2, 4, 8):
    os.environ["SID_BANDS"] = str(nb)
    timeit("run_pair, %d band(s), pinned" % nb, lambda: ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0))
os.environ["SID_BANDS"] = "8"
timeit("run_pair, 8 bands, pageable images", lambda: ctx.run_pair(img1, img2, c1, r1, c2, r2, b, s, angles, 0.0))
# host-side cost of the call without GPU work: tiny point set
timeit("run (resident) 64 points", lambda: ctx.run(c1[:64], r1[:64], c2[:64], r2[:64], b[:64], s, angles, 0.0))
