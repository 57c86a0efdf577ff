#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2c_pytest.txt; tail -15 gpurun_out/r2c_pytest.txt
timeout 400 python scratch/time_bands.py > gpurun_out/r2c_bands.txt 2>&1; cat gpurun_out/r2c_bands.txt
timeout 600 python bench.py 2> gpurun_out/bench_r02_d.err | tail -1 > gpurun_out/bench_r02_d.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_d.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['e2e_pageable'], d['roofline']['kernel_ms'], d['parity'])"
