#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_pytest18.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads entities" \
  "pf1|-DCCU_BVH_PREFETCH=1||--workloads entities" \
  "pf2|-DCCU_BVH_PREFETCH=2||--workloads entities" \
  "pf4|-DCCU_BVH_PREFETCH=4||--workloads entities" \
  "pf5|-DCCU_BVH_PREFETCH=5||--workloads entities" \
  "pf7|-DCCU_BVH_PREFETCH=7||--workloads entities" \
  "w32_r64|-DCCU_Q_WARPS=32||--workloads entities" \
  "w24|-DCCU_Q_WARPS=24||--workloads entities" \
  "pf5_stack8|-DCCU_BVH_PREFETCH=5 -DCCU_Q_STACK=8||--workloads entities" \
  "pf5_r6|-DCCU_BVH_PREFETCH=5|CCU_Q_REFILL_MIN=6|--workloads entities" \
  "pf5_r10|-DCCU_BVH_PREFETCH=5|CCU_Q_REFILL_MIN=10|--workloads entities"
