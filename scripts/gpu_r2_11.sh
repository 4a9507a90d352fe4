#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_pytest11.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities" \
  "sm8||CCU_Q_SHADE_MIN=8|--workloads config1,entities" \
  "sm12||CCU_Q_SHADE_MIN=12|--workloads config1" \
  "sm16||CCU_Q_SHADE_MIN=16|--workloads config1,entities" \
  "sm20||CCU_Q_SHADE_MIN=20|--workloads config1" \
  "sm24||CCU_Q_SHADE_MIN=24|--workloads config1,entities" \
  "sm28||CCU_Q_SHADE_MIN=28|--workloads config1" \
  "rows40_sm16|-DCCU_Q_ROWS=40|CCU_Q_SHADE_MIN=16|--workloads config1" \
  "rows40_sm24|-DCCU_Q_ROWS=40|CCU_Q_SHADE_MIN=24|--workloads config1" \
  "rows48_sm24|-DCCU_Q_ROWS=48|CCU_Q_SHADE_MIN=24|--workloads config1"
