#!/bin/bash
# fresh captures of the BVH kernel and the first-hit pass with the final code
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 1 -c 1 -o gpurun_out/r02d_queue_ent_prof -f \
   python scripts/qbench.py --workloads entities --reps 1 --passes 2 > gpurun_out/r02d_queue_ent_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_first_hit -s 1 -c 1 -o gpurun_out/r02d_first_hit_prof -f \
   python scripts/qbench.py --workloads config1 --reps 1 > gpurun_out/r02d_first_hit_ncu.log 2>&1
ls -la gpurun_out/r02d*.ncu-rep
