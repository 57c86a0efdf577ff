#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/time_variants.py cfg2 3000 40 2>&1 | tee gpurun_out/variants_small.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -6 gpurun_out/pytest_gpu.txt
timeout 300 python scratch/time_variants.py cfg2 2>&1 | tee gpurun_out/variants_cfg2.txt
timeout 300 python scratch/time_variants.py cfg1 2>&1 | tee gpurun_out/variants_cfg1.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pm_points -s 3 -c 1 -o gpurun_out/prof_pm6 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full6.log 2>&1
