"""cProfile of the whole drop-in call at EW size (synthetic matches in generic position)."""
import os, sys, time, io, cProfile, pstats, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sea_ice_drift_b200 import synthetic as syn, pmlib, _lib
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=0)
rng = np.random.default_rng(7); side = img1.shape[0]; nk = 50000
m = syn.rotation_matrix(img1.shape, 2.0)
kx, ky = rng.uniform(40, side - 40, nk), rng.uniform(40, side - 40, nk)
k2x, k2y = syn.apply_affine(m, kx, ky); k2x, k2y = k2x + rng.normal(0, 0.8, nk), k2y + rng.normal(0, 0.8, nk)
n1, n2 = syn.ArrayDomain(img1), syn.ArrayDomain(img2)
gx, gy = np.meshgrid(np.linspace(150, side - 150, 200), np.linspace(150, side - 150, 200))
lon, lat = n2.transform_points(gx, gy)
kw = dict(angles=[-3, 0, 3], img_size=35)
for _ in range(3):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res = pmlib.pattern_matching(lon, lat, n1, kx, ky, n2, k2x, k2y, **kw)
    print("call ms", round((time.perf_counter() - t0) * 1e3, 2))
pr = cProfile.Profile(); pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    res = pmlib.pattern_matching(lon, lat, n1, kx, ky, n2, k2x, k2y, **kw)
pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:4000])
ctx = _lib.default_context()
print("kernel", ctx.last_kernel_name, ctx.last_kernel_ms)
c2fg, r2fg, brd = pmlib.prepare_first_guess(np.round(gx.ravel()), np.round(gy.ravel()), n1, kx, ky, n2, k2x, k2y, 35)
print("borders", np.unique(brd, return_counts=True))
