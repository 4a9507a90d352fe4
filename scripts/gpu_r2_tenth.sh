#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1" \
  "fair13|-DCCU_Q_FAIR=13||--workloads config1" \
  "fair16|-DCCU_Q_FAIR=16||--workloads config1" \
  "rows40|-DCCU_Q_ROWS=40||--workloads config1" \
  "rows40_y0|-DCCU_Q_ROWS=40|CCU_YIELD_BELOW=0|--workloads config1" \
  "rows40_y0_b32|-DCCU_Q_ROWS=40|CCU_YIELD_BELOW=0 CCU_Q_MARCH_BIAS=32|--workloads config1" \
  "rows40_y8_b16|-DCCU_Q_ROWS=40|CCU_YIELD_BELOW=8 CCU_Q_MARCH_BIAS=16|--workloads config1" \
  "rows48_y0_b32|-DCCU_Q_ROWS=48|CCU_YIELD_BELOW=0 CCU_Q_MARCH_BIAS=32|--workloads config1" \
  "rows48_y0_b32_mw20|-DCCU_Q_ROWS=48|CCU_YIELD_BELOW=0 CCU_Q_MARCH_BIAS=32 CCU_Q_MARCH_WARPS=20|--workloads config1" \
  "rows32_y0_b32||CCU_YIELD_BELOW=0 CCU_Q_MARCH_BIAS=32|--workloads config1" \
  "base256|||--workloads config1 --passes 256 --reps 2" \
  "fair13_256|-DCCU_Q_FAIR=13||--workloads config1 --passes 256 --reps 2"
CHUNKYCU_LIB=/tmp/ccu_variants/rows40.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 2 -c 1 -o gpurun_out/r02_rows40_prof -f \
   python scripts/qbench.py --workloads config1 --reps 2 > gpurun_out/r02_rows40_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
