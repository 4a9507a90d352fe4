#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_acceptance.py -q -m gpu -x --timeout=600 -k "first_hit or preview or acceptance" 2>&1 | tail -4 | tee gpurun_out/r2_pytest21.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities,large" \
  "noskip|-DCCU_FH_WARP_SKIP=0||--workloads config1,large" \
  "fh3|-DCCU_FH_MIN_BLOCKS=3||--workloads config1" \
  "fh5|-DCCU_FH_MIN_BLOCKS=5||--workloads config1" \
  "fh6|-DCCU_FH_MIN_BLOCKS=6||--workloads config1"
