#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_pytest13.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities" \
  "fence0|-DCCU_FENCE_MODE=0||--workloads config1" \
  "fence2|-DCCU_FENCE_MODE=2||--workloads config1,entities" \
  "refill4||CCU_Q_REFILL_MIN=4|--workloads config1" \
  "refill6||CCU_Q_REFILL_MIN=6|--workloads config1" \
  "refill10||CCU_Q_REFILL_MIN=10|--workloads config1" \
  "refill12||CCU_Q_REFILL_MIN=12|--workloads config1" \
  "refill16||CCU_Q_REFILL_MIN=16|--workloads config1" \
  "yield12||CCU_YIELD_BELOW=12|--workloads config1" \
  "yield16||CCU_YIELD_BELOW=16|--workloads config1" \
  "yield24||CCU_YIELD_BELOW=24|--workloads config1" \
  "mw20||CCU_Q_MARCH_WARPS=20|--workloads config1" \
  "mw24||CCU_Q_MARCH_WARPS=24|--workloads config1" \
  "mw28||CCU_Q_MARCH_WARPS=28|--workloads config1" \
  "bias0||CCU_Q_MARCH_BIAS=0|--workloads config1" \
  "bias8||CCU_Q_MARCH_BIAS=8|--workloads config1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 2 -c 1 -o gpurun_out/r02b_queue_prof -f \
   python scripts/qbench.py --workloads config1 --reps 2 > gpurun_out/r02b_queue_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
