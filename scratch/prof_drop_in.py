"""cProfile of the whole drop-in call on BASELINE configs[0] (needs gpurun_out/matches_cfg1.npz or computes ORB matches)."""
import os, sys, time, io, cProfile, pstats, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from sea_ice_drift_b200 import synthetic as syn, pmlib
img1, img2, n1, n2, lon, lat = bench.drop_in_scene(syn)
x1, y1, x2, y2 = syn.orb_matches(img1, img2)
print("matches", len(x1), "integer keypoints:", float(np.mean(x1 == np.round(x1))))
kw = dict(angles=[0], img_size=35)
for fg in ("auto", "device", "host"):
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            res = pmlib.pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2, first_guess=fg, **kw)
        ts.append(time.perf_counter() - t0)
    print("pattern_matching first_guess=%-6s: %s ms" % (fg, [round(t * 1e3, 1) for t in ts]))
pr = cProfile.Profile(); pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    res = pmlib.pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2, **kw)
pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4500])
