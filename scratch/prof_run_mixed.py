"""One resident step on the EW pair with prepare_first_guess's mixed borders (for ncu): python scratch/prof_run_mixed.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sea_ice_drift_b200 import _lib, synthetic as syn
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=0)
rng = np.random.default_rng(7); side = img1.shape[0]; nk = 50000
m = syn.rotation_matrix(img1.shape, 2.0)
kx, ky = rng.uniform(40, side - 40, nk), rng.uniform(40, side - 40, nk)
k2x, k2y = syn.apply_affine(m, kx, ky); k2x, k2y = k2x + rng.normal(0, 0.8, nk), k2y + rng.normal(0, 0.8, nk)
pts = list(syn.orb_first_guess_inputs(img1, img2, 200, 35, inset=150, matches=(kx, ky, k2x, k2y)))
ctx = _lib.Context(0); ctx.set_pair(img1, img2)
for _ in range(3):
    out = ctx.run(*pts, 35, [-3, 0, 3], 0.0)
    print("kernel ms", ctx.last_kernel_ms, ctx.last_kernel_name, "points", len(pts[0]), "borders > 20:", int((pts[4] > 20).sum()))
