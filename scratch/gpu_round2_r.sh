#!/bin/bash
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2961$1 scratch/time_series.py --pairs 16 --distinct 1 --check 0 --gather $2 2>&1 | tail -1; }
run 8 root | tee gpurun_out/r2r_series_n8_root.txt
run 8 all | tee gpurun_out/r2r_series_n8_all.txt
run 4 all | tee gpurun_out/r2r_series_n4_all.txt
run 2 all | tee gpurun_out/r2r_series_n2_all.txt
