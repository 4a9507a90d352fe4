#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_pytest17.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities,indoor" \
  "noshort|-DCCU_SHADOW_SHORTCUT=0||--workloads entities" \
  "bm12||CCU_Q_BATCH_MIN=12|--workloads config1,entities" \
  "bm16||CCU_Q_BATCH_MIN=16|--workloads config1,entities,indoor" \
  "bm20||CCU_Q_BATCH_MIN=20|--workloads config1,entities,indoor" \
  "bm24||CCU_Q_BATCH_MIN=24|--workloads config1,entities" \
  "bm28||CCU_Q_BATCH_MIN=28|--workloads config1" \
  "bm20_r8||CCU_Q_BATCH_MIN=20 CCU_Q_BATCH_ROUNDS=8|--workloads config1" \
  "bm20_r2||CCU_Q_BATCH_MIN=20 CCU_Q_BATCH_ROUNDS=2|--workloads config1" \
  "bm20_ns500||CCU_Q_BATCH_MIN=20 CCU_Q_BATCH_NS=500|--workloads config1" \
  "bm20_mw22||CCU_Q_BATCH_MIN=20 CCU_Q_MARCH_WARPS=22|--workloads config1" \
  "bm24_mw24||CCU_Q_BATCH_MIN=24 CCU_Q_MARCH_WARPS=24|--workloads config1" \
  "bm20_mw28||CCU_Q_BATCH_MIN=20 CCU_Q_MARCH_WARPS=28|--workloads config1"
