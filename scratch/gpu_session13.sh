#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; cat gpurun_out/smoke.txt | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1_err.txt; echo "rc=$?" >> gpurun_out/bench_n1_err.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_err.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2_err.txt; echo "rc=$?" >> gpurun_out/bench_n2_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pm_points -s 3 -c 1 -o gpurun_out/prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
python -c "
import json
for f in ('bench_n1','bench_n2','bench_ref_n1'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d.get('clocks'), (d.get('cpu_baseline') or {}).get('value'))
"
