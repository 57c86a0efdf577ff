#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "run_pair or pageable or drop_in or border_classes or uploaded" 2>&1 | tail -4 > gpurun_out/r2v_pytest.txt; cat gpurun_out/r2v_pytest.txt
timeout 400 python scratch/time_bands2.py 2>&1 | grep -i "pageable\|default\|cores" > gpurun_out/r2v_bands.txt; cat gpurun_out/r2v_bands.txt
