"""Deformation step at PM-grid scale: sid_deformation (host arrays in/out) vs the NumPy restatement of the
reference's libdefor on the same triangulation."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sea_ice_drift_b200 import libdefor, _lib
from oracle import defor_oracle

rng = np.random.default_rng(0)
for n in (40000, 90000, 160000):
    x = rng.uniform(-4e5, 4e5, n); y = rng.uniform(-4e5, 4e5, n)
    u = rng.normal(0, 0.1, n); v = rng.normal(0, 0.1, n)
    t0 = time.perf_counter(); t = libdefor.triangulate(x, y); t_tri = time.perf_counter() - t0
    ctx = _lib.default_context()
    ctx.deformation(x, y, u, v, t)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); got = ctx.deformation(x, y, u, v, t); ts.append(time.perf_counter() - t0)
    tc = []
    for _ in range(5):
        t0 = time.perf_counter(); want = defor_oracle.deformation(x, y, u, v, t); tc.append(time.perf_counter() - t0)
    same = int(np.sum((got[3] == want[3]) & (got[4] == want[4])))
    print("nodes %6d elements %6d | Delaunay (host, scipy) %.1f ms | GPU incl. copies %.3f ms | NumPy %.2f ms | "
          "area+perimeter bit-identical %d / %d, max rel diff %.1e"
          % (n, len(t), 1e3 * t_tri, 1e3 * min(ts), 1e3 * min(tc), same, len(t), np.max(np.abs(got[3] / want[3] - 1))))
