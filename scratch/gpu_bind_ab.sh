#!/bin/bash
# N = 8 bench with and without binding each rank to its GPU's NUMA node
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo.txt
for mode in bind nobind; do
  if [ $mode = nobind ]; then export SID_NO_BIND=1; else unset SID_NO_BIND; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 2>gpurun_out/bench_n8_$mode.err | tail -1 > gpurun_out/bench_n8_$mode.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_n8_$mode.json')); print('$mode', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config'].get('host_binding'))"
done
