#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_pytest12.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities,large" \
  "noflat|-DCCU_FLAT_MARCH=0||--workloads config1,entities" \
  "notail|-DCCU_SHADE_ONE_TAIL=0||--workloads config1" \
  "sticky8||CCU_Q_STICKY=8|--workloads config1,entities" \
  "sticky16||CCU_Q_STICKY=16|--workloads config1,entities" \
  "sticky24||CCU_Q_STICKY=24|--workloads config1" \
  "fence1|-DCCU_FENCE_MODE=1||--workloads config1" \
  "fence2|-DCCU_FENCE_MODE=2||--workloads config1" \
  "fence2_sticky16|-DCCU_FENCE_MODE=2|CCU_Q_STICKY=16|--workloads config1"
