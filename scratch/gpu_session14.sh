#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pm_ -s 6 -c 4 --csv --log-file gpurun_out/split_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
grep -v "^==" gpurun_out/split_launches.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[1:]:
    if len(r)>10: print(r[h.index('Kernel Name')][:40], r[h.index('Metric Name')], r[h.index('Metric Value')])
"
