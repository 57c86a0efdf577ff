#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "first_guess_on_device_ew or run_pair" -s 2>&1 | tail -30 > gpurun_out/r2b_pytest.txt; tail -30 gpurun_out/r2b_pytest.txt
timeout 400 python scratch/time_bands.py > gpurun_out/r2b_bands.txt 2>&1; cat gpurun_out/r2b_bands.txt
SID_LIBRARY=$PWD/scratch/ab/libsid_wsprof.so timeout 200 python scratch/ws_prof.py > gpurun_out/r2b_wsprof.txt 2>&1; cat gpurun_out/r2b_wsprof.txt
