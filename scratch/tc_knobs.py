"""cfg2/cfg3 kernel time for combinations of SID_TC_NACC / SID_TC_SPW (and checks the outputs stay identical)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sea_ice_drift_b200 import _lib, synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(name, seed=0)
ctx = _lib.Context(0)
ctx.set_pair(img1, img2)
ref = None
for nacc, spw in ((3, 1), (2, 2), (1, 2), (1, 3), (2, 1), (1, 1)):
    os.environ["SID_TC_NACC"] = str(nacc); os.environ["SID_TC_SPW"] = str(spw)
    ms = []
    for _ in range(3):
        out = ctx.run(c1, r1, c2, r2, b, cfg["img_size"], cfg["angles"], 0.0)
        ms.append(ctx.last_kernel_ms)
    if ref is None:
        ref = out
    print("%s nacc=%d spw=%d: %.3f ms  same=%s" % (name, nacc, spw, min(ms), np.array_equal(out, ref, equal_nan=True)), flush=True)
