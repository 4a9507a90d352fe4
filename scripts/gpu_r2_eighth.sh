#!/bin/bash
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -5 | tee gpurun_out/r2_pytest8.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "pred|||--workloads entities,config1" \
  "pred_rows24|-DCCU_Q_ROWS=24||--workloads entities,config1" \
  "pred_rows28|-DCCU_Q_ROWS=28||--workloads entities" \
  "pred_s10|-DCCU_Q_STACK=10||--workloads entities"
