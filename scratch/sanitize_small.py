import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sea_ice_drift_b200 import _lib, synthetic as syn
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=1, side=700, grid=8)
b = np.floor(np.random.default_rng(0).uniform(20, 30, b.size))
ctx = _lib.Context(0)
out = ctx.run_pair(img1, img2, c1, r1, c2, r2, b, 35, [-3, 0, 3], 0.0)
# search radius <= 20 (s = 35; the shared-memory budget decides): the warp-specialised pipeline kernel (pm_ws_kernel), 3 and 7 angles, mixed borders, single angle
b20 = np.floor(np.random.default_rng(3).uniform(8, 21, b.size))
out_ws = ctx.run(c1, r1, c2, r2, b20, 35, [-3, 0, 3], 0.0); name_ws = ctx.last_kernel_name
out_ws7 = ctx.run(c1, r1, c2, r2, b20, 34, list(range(-3, 4)), 0.0, rot_order=1, flags=7)
out_ws1 = ctx.run(c1, r1, c2, r2, b20, 21, [0], 0.0)
assert name_ws == "sid::pm_ws_kernel", name_ws
print("ws", name_ws, np.isnan(out_ws[:, 0]).sum(), np.isnan(out_ws7[:, 0]).sum(), np.isnan(out_ws1[:, 0]).sum())
out2 = ctx.run(c1, r1, c2, r2, b, 35, list(range(-3, 4)), 0.0, rot_order=1, flags=7)
out3 = ctx.run(c1[:20], r1[:20], c2[:20], r2[:20], b[:20] + 40, 50, [0, 2], 0.0)
print("ok", np.isnan(out[:, 0]).sum(), np.isnan(out2[:, 0]).sum(), np.isnan(out3[:, 0]).sum())
# multi-band pageable upload (staged through the pinned double buffer), matcher, deformation, single-call entry points
img1b, img2b, d1, e1, d2, e2, bb, _ = syn.make_config("cfg2", seed=2, side=1300, grid=6)
out4 = ctx.run_pair(img1b[:1100], img2b, d1, e1, d2, e2, bb, 35, [-3, 0, 3], 0.0)
rng = np.random.default_rng(1)
q = rng.integers(0, 256, (700, 32), dtype=np.uint8); t = rng.integers(0, 256, (900, 32), dtype=np.uint8)
idx, dist = ctx.knn_hamming2(q, t)
x = rng.uniform(0, 1e5, 500); y = rng.uniform(0, 1e5, 500)
from sea_ice_drift_b200 import libdefor
tri = libdefor.triangulate(x, y)
e = ctx.deformation(x, y, rng.normal(0, .1, 500), rng.normal(0, .1, 500), tri)
tpl = ctx.get_template(img1, 300.5, 310.25, 2.0, 35)
res = ctx.match_template(img2[200:300, 220:330], tpl)
hes = ctx.get_hessian(res)
print("ok2", np.isnan(out4[:, 0]).sum(), idx.shape, len(tri), tpl.shape, res.shape, float(hes.max()))
# border classes (two launches per band, tail regions with one stride) through the banded, staged upload of pageable images,
# and the two-statistics-set layout of pm_ws_kernel (borders 21..22)
import os
os.environ["SID_BANDS"] = "3"
bmix = np.where(np.random.default_rng(5).random(b.size) < 0.8, 20.0, np.floor(np.random.default_rng(6).uniform(21, 40, b.size)))
c1x, r1x, c2x, r2x = [np.tile(v, 8) for v in (c1, r1, c2, r2)]; bmx = np.tile(bmix, 8)
out_mix = ctx.run_pair(img1, img2, c1x, r1x, c2x, r2x, bmx, 35, [-3, 0, 3], 0.0)
del os.environ["SID_BANDS"]
out_mix1 = ctx.run(c1x, r1x, c2x, r2x, bmx, 35, [-3, 0, 3], 0.0); name_mix = ctx.last_kernel_name
assert np.array_equal(out_mix, out_mix1, equal_nan=True)
b22 = np.floor(np.random.default_rng(7).uniform(21, 23, b.size))
out_22 = ctx.run(c1, r1, c2, r2, b22, 35, [-3, 0, 3], 0.0); name_22 = ctx.last_kernel_name
print("classes", name_mix, len(c1x), int((bmx <= 22).sum()), np.isnan(out_mix[:, 0]).sum(), "| two statistics sets", name_22, np.isnan(out_22[:, 0]).sum())
assert "+ sid::pm_ws_kernel" in name_mix and name_22 == "sid::pm_ws_kernel", (name_mix, name_22)
