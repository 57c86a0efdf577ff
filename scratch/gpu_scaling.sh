#!/bin/bash
# bench.py at N = 2, 4, 8 on one 8-GPU box (the driver's scaling run), JSON lines into gpurun_out/
mkdir -p gpurun_out
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n 2>gpurun_out/bench_n$n.err | tail -1 > gpurun_out/bench_n$n.json
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n$n.json')); print($n, d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
done
