#!/bin/bash
# 2-GPU check: NCCL tests, bench at N=2 (weak + strong block), split timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "nccl or sharded or two_gpus or template_matcher or single_call or series" 2>&1 | tail -4 > gpurun_out/r2f_pytest.txt; cat gpurun_out/r2f_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 2>gpurun_out/bench_r02_n2.err | tail -1 > gpurun_out/bench_r02_n2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_n2.json')); print(2, d['value'], d['e2e'], d['ms_per_step'], d['strong'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 scratch/split_timing.py > gpurun_out/r2f_split_n2.txt 2>&1; tail -5 gpurun_out/r2f_split_n2.txt
