#!/bin/bash
# round-2 final checkpoint on one B200: GPU suite, smoke, bench (both arms), then the evidence script
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.txt
timeout 900 python bench.py --steps 100 --warmup 5 2> gpurun_out/bench_r02_e.err | tail -1 > gpurun_out/bench_r02_e.json; cut -c1-900 gpurun_out/bench_r02_e.json
timeout 900 python bench.py --impl reference 2> gpurun_out/bench_r02_ref.err | tail -1 > gpurun_out/bench_r02_ref.json; cut -c1-400 gpurun_out/bench_r02_ref.json
bash scratch/gpu_evidence_ws.sh
