#!/bin/bash
# round-2 evidence for pm_ws_kernel + pm_tail_kernel on one B200: sanitizer (memcheck; racecheck per kernel), ncu --set full
# (both kernels of a step), launch list of the bench
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scratch/sanitize_small.py > gpurun_out/ws_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/ws_memcheck.txt; tail -4 gpurun_out/ws_memcheck.txt
timeout 500 compute-sanitizer --tool racecheck --kernel-regex kns=pm_tail --error-exitcode 9 python scratch/sanitize_small.py > gpurun_out/tail_racecheck.txt 2>&1; echo "racecheck(pm_tail) rc=$?" >> gpurun_out/tail_racecheck.txt; tail -4 gpurun_out/tail_racecheck.txt
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=pm_ws --error-exitcode 9 python scratch/sanitize_small.py > gpurun_out/ws_racecheck.txt 2>&1; echo "racecheck(pm_ws) rc=$?" >> gpurun_out/ws_racecheck.txt; tail -6 gpurun_out/ws_racecheck.txt
timeout 500 ncu --set full --import-source on --clock-control none -k regex:'pm_ws|pm_tail' --launch-skip 2 --launch-count 2 -f -o gpurun_out/ws_final python scratch/prof_run.py > gpurun_out/ws_final_ncu.log 2>&1; tail -2 gpurun_out/ws_final_ncu.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_cfg2_ws.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu_ws.log 2>&1; tail -c 300 gpurun_out/r02_bench_under_ncu_ws.log
