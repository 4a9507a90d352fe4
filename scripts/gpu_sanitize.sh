#!/bin/bash
# compute-sanitizer over a small render of every scene kind (default kernel + first-hit pass): memcheck + initcheck (+ racecheck with RACE=1)
mkdir -p gpurun_out
tools="memcheck initcheck"
[ -n "$RACE" ] && tools="$tools racecheck"
for tool in $tools; do
  for sc in terrain256 entities; do
    echo "== $tool $sc"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/run_render.py --scene $sc --width 256 --height 144 --passes 2 --windows 1 --kernel 4 --first-hit 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|window 0|first hit" | head -8
  done
done
