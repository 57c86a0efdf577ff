"""Same-box A/B of library builds: python scratch/ab_libs.py name1=path1.so name2=path2.so [workload ...]
Runs each library in its own subprocess, alternating, several rounds; prints the best and median kernel time per library."""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np
from sea_ice_drift_b200 import _lib, synthetic as syn
name = sys.argv[1]
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(name, seed=0)
ctx = _lib.Context(0); ctx.set_pair(img1, img2)
ms = []
for _ in range(8):
    out = ctx.run(c1, r1, c2, r2, b, cfg["img_size"], cfg["angles"], 0.0)
    ms.append(ctx.last_kernel_ms)
print(json.dumps({"ms": ms[2:], "sum": float(np.nansum(out))}))
''' % ROOT
libs = [a.split("=", 1) for a in sys.argv[1:] if "=" in a]
loads = [a for a in sys.argv[1:] if "=" not in a] or ["cfg2"]
for wl in loads:
    res = {n: [] for n, _ in libs}
    sums = {}
    for rnd in range(3):
        for n, path in libs:
            env = dict(os.environ, SID_LIBRARY=os.path.join(ROOT, path))
            out = subprocess.run([sys.executable, "-c", CHILD, wl], env=env, capture_output=True, text=True)
            d = json.loads(out.stdout.strip().splitlines()[-1])
            res[n] += d["ms"]; sums[n] = d["sum"]
    for n, _ in libs:
        print("%s %-14s best %.3f ms  median %.3f ms  checksum %.6f" % (wl, n, min(res[n]), float(np.median(res[n])), sums[n]), flush=True)
