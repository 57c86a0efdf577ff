#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
timeout 300 python scratch/time_variants.py cfg2 2>&1 | head -3 | tee gpurun_out/variants_cfg2.txt
timeout 300 python scratch/time_variants.py cfg3 2>&1 | head -3 | tee gpurun_out/variants_cfg3.txt
timeout 600 python scratch/time_variants.py cfg4 6000 120 2>&1 | head -3 | tee gpurun_out/variants_cfg4.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pm_points -s 3 -c 1 -o gpurun_out/prof_pm7 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full7.log 2>&1
