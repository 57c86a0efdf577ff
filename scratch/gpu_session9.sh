#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
python scratch/time_variants.py cfg2 2>&1 | head -3 | tee gpurun_out/variants_cfg2.txt
ncu --set full --clock-control none --import-source on -k regex:pm_points -s 3 -c 1 -o gpurun_out/prof_pm5 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full5.log 2>&1
