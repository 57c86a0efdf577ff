#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
SID_PM_PATH=dp4a python -m pytest tests -m gpu -x -q -k "seeded or golden_reference" > gpurun_out/pytest_gpu_dp4a.txt 2>&1; tail -3 gpurun_out/pytest_gpu_dp4a.txt
python scratch/time_variants.py cfg2 2>&1 | tee gpurun_out/variants_cfg2.txt
python scratch/time_variants.py cfg1 2>&1 | tee gpurun_out/variants_cfg1.txt
python scratch/time_variants.py cfg4 6000 120 2>&1 | tee gpurun_out/variants_cfg4.txt
python scratch/time_variants.py cfg3 5000 60 2>&1 | tee gpurun_out/variants_cfg3.txt
