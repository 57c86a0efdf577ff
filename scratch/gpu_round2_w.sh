#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2w_pytest.txt; tail -5 gpurun_out/r2w_pytest.txt
timeout 600 python scratch/time_classes.py 2>&1 | head -5 > gpurun_out/r2w_classes.txt; cat gpurun_out/r2w_classes.txt
timeout 300 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from sea_ice_drift_b200 import _lib, synthetic as syn
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=0)
ctx = _lib.Context(0); ctx.set_pair(img1, img2)
for brd in (20, 21, 22, 23):
    bb = np.full(len(c1), float(brd))
    for _ in range(3): out = ctx.run(c1, r1, c2, r2, bb, 35, [-3, 0, 3], 0.0)
    print("uniform border", brd, ctx.last_kernel_name, round(ctx.last_kernel_ms, 3), "ms", flush=True)
PY
