#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py 2> gpurun_out/bench_r02_g.err | tail -1 > gpurun_out/bench_r02_g.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_g.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']); print(json.dumps(d['drop_in'])[:1500]); print(json.dumps(d['drop_in_ew'])); print(json.dumps(d['configs']['cfg1'])[:400])"; tail -5 gpurun_out/bench_r02_g.err
timeout 600 python -m pytest tests/test_bench_contract.py -q -m gpu 2>&1 | tail -3
