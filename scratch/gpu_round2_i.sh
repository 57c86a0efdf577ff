#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 2>gpurun_out/bench_r02_n8.err | tail -1 > gpurun_out/bench_r02_n8.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_n8.json')); print(8, d['value'], d['e2e'], d['ms_per_step'], d['strong'], d['clocks'])"
