"""Role profile of pm_ws_kernel (library built with -DSID_WS_PROF; SID_LIBRARY points at it): cycles per point per bucket."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SID_PM_PATH"] = "ws"
from sea_ice_drift_b200 import _lib, synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(name, seed=0)
ctx = _lib.Context(0); ctx.set_pair(img1, img2)
for _ in range(2):
    out = ctx.run(c1, r1, c2, r2, b, cfg["img_size"], cfg["angles"], 0.0)
    print("kernel ms", ctx.last_kernel_ms, flush=True)
