import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sea_ice_drift_b200 import _lib, synthetic as syn
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=1, side=700, grid=8)
b = np.floor(np.random.default_rng(0).uniform(20, 30, b.size))
ctx = _lib.Context(0)
out = ctx.run_pair(img1, img2, c1, r1, c2, r2, b, 35, [-3, 0, 3], 0.0)
out2 = ctx.run(c1, r1, c2, r2, b, 35, list(range(-3, 4)), 0.0, rot_order=1, flags=7)
out3 = ctx.run(c1[:20], r1[:20], c2[:20], r2[:20], b[:20] + 40, 50, [0, 2], 0.0)
print("ok", np.isnan(out[:, 0]).sum(), np.isnan(out2[:, 0]).sum(), np.isnan(out3[:, 0]).sum())
