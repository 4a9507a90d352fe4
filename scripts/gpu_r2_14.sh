#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_pytest14.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities,indoor" \
  "mw20||CCU_Q_MARCH_WARPS=20|--workloads config1,indoor" \
  "mw18||CCU_Q_MARCH_WARPS=18|--workloads config1" \
  "mw20_y16||CCU_Q_MARCH_WARPS=20 CCU_YIELD_BELOW=16|--workloads config1,indoor" \
  "mw20_y16_r6||CCU_Q_MARCH_WARPS=20 CCU_YIELD_BELOW=16 CCU_Q_REFILL_MIN=6|--workloads config1,indoor" \
  "mw20_y12_r6||CCU_Q_MARCH_WARPS=20 CCU_YIELD_BELOW=12 CCU_Q_REFILL_MIN=6|--workloads config1" \
  "mw20_b8||CCU_Q_MARCH_WARPS=20 CCU_Q_MARCH_BIAS=8|--workloads config1" \
  "nimath|-DCCU_NI_MATH||--workloads config1" \
  "nimat|-DCCU_NI_MATERIAL||--workloads config1" \
  "nimath_mw20|-DCCU_NI_MATH|CCU_Q_MARCH_WARPS=20|--workloads config1" \
  "w32|-DCCU_Q_WARPS=32|CCU_Q_MARCH_WARPS=24|--workloads config1" \
  "w24|-DCCU_Q_WARPS=24|CCU_Q_MARCH_WARPS=18|--workloads config1"
