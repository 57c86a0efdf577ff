"""prepare_first_guess at EW size: device path vs host path, and the pieces of the device path."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sea_ice_drift_b200 as sid
from sea_ice_drift_b200 import _lib, synthetic as syn, pmlib
from sea_ice_drift_b200.lib import interpolation_poly
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import _fg_case

def best(f, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts)

for side, n_kp, grid in ((2000, 15000, 50), (10400, 50000, 200), (10400, 50000, 400)):
    n1, n2, kx, ky, k2x, k2y, gx, gy = _fg_case(side, n_kp, grid, 3)
    ctx = _lib.default_context()
    f_dev = lambda: sid.prepare_first_guess(gx, gy, n1, kx, ky, n2, k2x, k2y, 35, first_guess='device')
    f_host = lambda: sid.prepare_first_guess(gx, gy, n1, kx, ky, n2, k2x, k2y, 35, first_guess='host')
    f_dev()
    t_dev, t_host = best(f_dev), best(f_host, 2)
    t_call = best(lambda: ctx.first_guess(kx, ky, k2x, k2y, np.uint16(k2x), np.uint16(k2y), gx, gy))
    t_poly = best(lambda: interpolation_poly(kx, ky, k2x, k2y, gx, gy))
    t_tp = best(lambda: n2.transform_points(*n1.transform_points(kx, ky), 1))
    print("%5d^2, %5d keypoints, %3dx%3d grid: prepare_first_guess device %.1f ms (sid_first_guess call %.1f, polynomial %.1f, transforms %.1f), host %.1f ms"
          % (side, n_kp, grid, grid, t_dev, t_call, t_poly, t_tp, t_host), flush=True)
