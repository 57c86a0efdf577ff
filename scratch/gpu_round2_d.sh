#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "warp_specialised or seeded or golden or randomised" 2>&1 | tail -5 > gpurun_out/r2d_pytest.txt; tail -5 gpurun_out/r2d_pytest.txt
timeout 900 python scratch/ab_libs.py head=scratch/ab/libsid_head.so new=sea_ice_drift_b200/libsid_b200.so cfg2 cfg3 > gpurun_out/r2d_ab.txt 2>&1; cat gpurun_out/r2d_ab.txt
