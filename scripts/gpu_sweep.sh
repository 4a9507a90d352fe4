#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -3 | tee gpurun_out/r2_pytest27.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads entities,config1" \
  "nopair|-DCCU_BVH_PAIR_LOADS=0||--workloads entities" \
  "pair_r6||CCU_Q_REFILL_MIN=6|--workloads entities" \
  "pair_r12||CCU_Q_REFILL_MIN=12|--workloads entities" \
  "pair_sticky8||CCU_Q_STICKY=8|--workloads entities"
