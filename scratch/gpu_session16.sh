#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n${N}_err.txt; echo "N=$N rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('N=%d value %.4g vec/s  ms/step %.3f  e2e %.4g  clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.json 2>/dev/null; head -c 200 gpurun_out/bench_ref_n8.json; echo
