#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=600 2>&1 | tail -3 | tee gpurun_out/r2_pytest23.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads entities,config1" \
  "pf4|-DCCU_BVH_PREFETCH=4||--workloads entities" \
  "stack10|-DCCU_Q_STACK=10||--workloads entities" \
  "stack14|-DCCU_Q_STACK=14||--workloads entities" \
  "bw24||CCU_Q_BVH_WARPS=24|--workloads entities" \
  "bw20||CCU_Q_BVH_WARPS=20|--workloads entities" \
  "sticky12||CCU_Q_STICKY=12|--workloads entities" \
  "sticky20||CCU_Q_STICKY=20|--workloads entities"
