#!/bin/bash
# usage: scripts/profile_kernel.sh <tag> <run_render args...>   -> gpurun_out/<tag>_prof.ncu-rep
TAG=$1; shift
mkdir -p gpurun_out
python scripts/run_render.py "$@" | tee gpurun_out/${TAG}_run.log
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/${TAG}_prof python scripts/run_render.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
