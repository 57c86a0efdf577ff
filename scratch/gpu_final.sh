#!/bin/bash
# round-end validation on one B200: full GPU suite, smoke, default bench (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.txt
timeout 900 python bench.py 2> gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json; cat gpurun_out/bench_n1.json
timeout 900 python bench.py --impl reference 2> gpurun_out/bench_ref.err | tail -1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
