#!/bin/bash
mkdir -p gpurun_out
python scratch/time_e2e.py 2>&1 | grep -E "run_pair|set_pair \+ run" | tee gpurun_out/time_e2e.txt
SID_BANDS=16 python -c "
import os,sys,time; sys.path.insert(0,'.')
import numpy as np, torch
from sea_ice_drift_b200 import _lib, synthetic as syn
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config('cfg2', seed=0)
p1=torch.from_numpy(img1).pin_memory().numpy(); p2=torch.from_numpy(img2).pin_memory().numpy()
ctx=_lib.Context(0)
for nb in (8,12,16):
    os.environ['SID_BANDS']=str(nb)
    for _ in range(2): ctx.run_pair(p1,p2,c1,r1,c2,r2,b,35,cfg['angles'],0.0)
    t=time.perf_counter()
    for _ in range(6): ctx.run_pair(p1,p2,c1,r1,c2,r2,b,35,cfg['angles'],0.0)
    print('bands',nb,'%.3f ms'%((time.perf_counter()-t)/6*1e3))
" 2>&1 | tee -a gpurun_out/time_e2e.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scratch/sharded_2gpu.py 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tee gpurun_out/sharded_2gpu.txt
