#!/bin/bash
# Template of a tuning batch on one B200 box: parity first, then same-box A/B timings (see scripts/gpu_variants.sh for the spec format).
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -3
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,indoor,entities,large" \
  "mw17||CCU_Q_MARCH_WARPS=17|--workloads config1,indoor" \
  "r10||CCU_Q_REFILL_MIN=10|--workloads config1,indoor"
