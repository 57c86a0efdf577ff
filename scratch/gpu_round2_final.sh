#!/bin/bash
# end-of-round validation on one B200 with the final library: GPU suite, smoke, bench (both arms), launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.txt
timeout 900 python bench.py --steps 100 --warmup 5 2> gpurun_out/bench_r02_final.err | tail -1 > gpurun_out/bench_r02_final.json; cut -c1-600 gpurun_out/bench_r02_final.json
timeout 900 python bench.py --impl reference 2> gpurun_out/bench_r02_ref.err | tail -1 > gpurun_out/bench_r02_ref.json; cut -c1-300 gpurun_out/bench_r02_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_cfg2_ws.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_under_ncu_ws.log 2>&1; tail -c 200 gpurun_out/r02_bench_under_ncu_ws.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scratch/sanitize_small.py > gpurun_out/ws_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/ws_memcheck.txt; tail -3 gpurun_out/ws_memcheck.txt
