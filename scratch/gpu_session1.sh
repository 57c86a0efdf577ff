#!/bin/bash
# first GPU session: parity tests, smoke, bench, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench_err.txt; echo "bench rc=$?" >> gpurun_out/bench_err.txt
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pm_points -s 3 -c 1 -o gpurun_out/prof_pm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt; cat gpurun_out/bench.json; tail -3 gpurun_out/bench_err.txt; cat gpurun_out/bench_ref.json
