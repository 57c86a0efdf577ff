#!/bin/bash
mkdir -p gpurun_out
for path in imma dp4a; do
  for tool in memcheck racecheck; do
    echo "== $path $tool"
    SID_PM_PATH=$path timeout 500 compute-sanitizer --tool $tool --print-limit 5 python scratch/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|^ok|=========     at" | head -8
  done
done
SID_PM_TMA=0 timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python scratch/sanitize_small.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard|^ok" | head -5
