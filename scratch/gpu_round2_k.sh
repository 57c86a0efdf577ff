#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2k_pytest.txt; tail -8 gpurun_out/r2k_pytest.txt
timeout 900 python bench.py 2> gpurun_out/bench_r02_f.err | tail -1 > gpurun_out/bench_r02_f.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_f.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['parity']['max_abs_dh']); print(json.dumps(d['configs']['cfg1']))"; tail -3 gpurun_out/bench_r02_f.err
