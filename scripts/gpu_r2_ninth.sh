#!/bin/bash
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=600 2>&1 | tail -5 | tee gpurun_out/r2_pytest9.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,indoor" \
  "rows40|-DCCU_Q_ROWS=40||--workloads config1,indoor" \
  "rows48|-DCCU_Q_ROWS=48||--workloads config1,indoor" \
  "rows64|-DCCU_Q_ROWS=64||--workloads config1" \
  "shade8||CCU_Q_SHADE_MIN=8|--workloads config1" \
  "shade16||CCU_Q_SHADE_MIN=16|--workloads config1" \
  "shade24||CCU_Q_SHADE_MIN=24|--workloads config1" \
  "rows48_w32|-DCCU_Q_ROWS=48 -DCCU_Q_WARPS=32||--workloads config1"
