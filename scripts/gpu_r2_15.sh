#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_pytest15.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities,indoor" \
  "mw18||CCU_Q_MARCH_WARPS=18|--workloads config1,indoor" \
  "mw16||CCU_Q_MARCH_WARPS=16|--workloads config1" \
  "mw18_y12_r6||CCU_Q_MARCH_WARPS=18 CCU_YIELD_BELOW=12 CCU_Q_REFILL_MIN=6|--workloads config1,indoor" \
  "w24_mw18|-DCCU_Q_WARPS=24|CCU_Q_MARCH_WARPS=18|--workloads config1,entities,indoor" \
  "w24_mw16|-DCCU_Q_WARPS=24|CCU_Q_MARCH_WARPS=16|--workloads config1,indoor" \
  "w24_mw20|-DCCU_Q_WARPS=24|CCU_Q_MARCH_WARPS=20|--workloads config1" \
  "w24_mw14|-DCCU_Q_WARPS=24|CCU_Q_MARCH_WARPS=14|--workloads config1" \
  "w20_mw14|-DCCU_Q_WARPS=20|CCU_Q_MARCH_WARPS=14|--workloads config1" \
  "w20_mw16|-DCCU_Q_WARPS=20|CCU_Q_MARCH_WARPS=16|--workloads config1" \
  "w24_mw18_y12_r6|-DCCU_Q_WARPS=24|CCU_Q_MARCH_WARPS=18 CCU_YIELD_BELOW=12 CCU_Q_REFILL_MIN=6|--workloads config1" \
  "w24_mw18_rows28|-DCCU_Q_WARPS=24 -DCCU_Q_ROWS=28|CCU_Q_MARCH_WARPS=18|--workloads config1"
