#!/bin/bash
# compute-sanitizer over a small render of every scene kind (default kernel): memcheck + racecheck + initcheck summaries
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  for sc in terrain256 entities; do
    echo "== $tool $sc"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/run_render.py --scene $sc --width 256 --height 144 --passes 2 --windows 1 --kernel 4 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|window 0" | head -8
  done
done
