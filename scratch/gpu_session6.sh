#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1_err.txt; echo "rc=$?" >> gpurun_out/bench_n1_err.txt
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2_err.txt; echo "rc=$?" >> gpurun_out/bench_n2_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2>> gpurun_out/bench_ref_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
cat gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1_err.txt; cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2_err.txt; cat gpurun_out/bench_ref_n2.json
