#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/variants.jsonl
NI="-DCCU_NI_MARCH_BEGIN -DCCU_NI_MATH -DCCU_NI_MATERIAL"
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,indoor" \
  "ni_mb|-DCCU_NI_MARCH_BEGIN||--workloads config1" \
  "ni_math|-DCCU_NI_MATH||--workloads config1" \
  "ni_mat|-DCCU_NI_MATERIAL||--workloads config1" \
  "ni_all|$NI||--workloads config1,indoor" \
  "ni_all_w32|$NI -DCCU_Q_WARPS=32||--workloads config1" \
  "ni_all_w24|$NI -DCCU_Q_WARPS=24||--workloads config1" \
  "w32|-DCCU_Q_WARPS=32||--workloads config1" \
  "base_again|||--workloads config1" \
  "yield16||CCU_YIELD_BELOW=16|--workloads config1" \
  "yield24||CCU_YIELD_BELOW=24|--workloads config1" \
  "yield28||CCU_YIELD_BELOW=28|--workloads config1" \
  "refill4||CCU_Q_REFILL_MIN=4|--workloads config1" \
  "refill12||CCU_Q_REFILL_MIN=12|--workloads config1" \
  "mw20||CCU_Q_MARCH_WARPS=20|--workloads config1" \
  "mw24||CCU_Q_MARCH_WARPS=24|--workloads config1" \
  "mw28||CCU_Q_MARCH_WARPS=28|--workloads config1" \
  "bias0||CCU_Q_MARCH_BIAS=0|--workloads config1" \
  "bias8||CCU_Q_MARCH_BIAS=8|--workloads config1" \
  "stats|-DCCU_Q_STATS||--workloads config1 --reps 1"
