#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1_err.txt; echo "rc=$?" >> gpurun_out/bench_n1_err.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value %.3g vec/s  ms/step %.3f | e2e %.3g vec/s  ms/step %.3f | parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity']))"
tail -2 gpurun_out/bench_n1_err.txt
