#!/bin/bash
# cfg5 time series at N = 8 (root gather), 4 and 2 (all-gather); run on an 8-GPU box
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29611 scratch/time_series.py --pairs 16 --distinct 1 --check 0 --gather $2 2>&1 | tail -1; }
run 8 root
run 4 all
run 2 all
