#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities,indoor" \
  "sm1|-DCCU_Q_STEAL_MARCH=1||--workloads config1,entities" \
  "ss1|-DCCU_Q_STEAL_SHADE=1||--workloads config1,entities" \
  "sm1_ss1|-DCCU_Q_STEAL_MARCH=1 -DCCU_Q_STEAL_SHADE=1||--workloads config1,entities,indoor" \
  "sm2_ss2|-DCCU_Q_STEAL_MARCH=2 -DCCU_Q_STEAL_SHADE=2||--workloads config1,entities,indoor" \
  "sm1_ss2|-DCCU_Q_STEAL_MARCH=1 -DCCU_Q_STEAL_SHADE=2||--workloads config1" \
  "sm2_ss1|-DCCU_Q_STEAL_MARCH=2 -DCCU_Q_STEAL_SHADE=1||--workloads config1"
CHUNKYCU_LIB=/tmp/ccu_variants/sm1_ss1.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=600 2>&1 | tail -3 | tee gpurun_out/r2_pytest22.log
