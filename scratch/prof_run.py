"""One cfg2-shaped launch for ncu (usage: python scratch/prof_run.py [side] [grid] [cfg])."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sea_ice_drift_b200 import _lib, synthetic as syn

side = int(sys.argv[1]) if len(sys.argv) > 1 else 10400
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 200
name = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(name, seed=0, side=side, grid=grid)
ctx = _lib.Context(0)
ctx.set_pair(img1, img2)
for _ in range(3):
    out = ctx.run(c1, r1, c2, r2, b, cfg["img_size"], cfg["angles"], 0.0)
    print("kernel ms", ctx.last_kernel_ms, "points", len(c1))
