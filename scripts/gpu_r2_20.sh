#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,large" \
  "rows36|-DCCU_Q_ROWS=36||--workloads config1" \
  "rows40|-DCCU_Q_ROWS=40||--workloads config1" \
  "rows28|-DCCU_Q_ROWS=28||--workloads config1" \
  "w26|-DCCU_Q_WARPS=26|CCU_Q_MARCH_WARPS=17|--workloads config1" \
  "fh3|-DCCU_FH_MIN_BLOCKS=3||--workloads config1" \
  "fh5|-DCCU_FH_MIN_BLOCKS=5||--workloads config1" \
  "fh6|-DCCU_FH_MIN_BLOCKS=6||--workloads config1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_first_hit -s 1 -c 1 -o gpurun_out/r02c_first_hit_prof -f \
   env CHUNKYCU_LIB=/tmp/ccu_variants/base.so python scripts/qbench.py --workloads config1 --reps 1 > gpurun_out/r02c_first_hit_ncu.log 2>&1
ls -la gpurun_out/r02c_first_hit_prof.ncu-rep
