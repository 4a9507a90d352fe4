#!/bin/bash
# ncu evidence for the render kernel.  usage: scripts/profile.sh <tag> [run_render.py args]
TAG=$1; shift
mkdir -p gpurun_out
python scripts/run_render.py "$@" | tee gpurun_out/${TAG}_run.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/run_render.py "$@" > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/${TAG}_prof python scripts/run_render.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/
