#!/bin/bash
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
timeout 1500 python -m pytest tests -q -m gpu -x --timeout=900 2>&1 | tail -12 | tee gpurun_out/r2_pytest4.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,indoor,entities,large" \
  "fh2|-DCCU_FH_MIN_BLOCKS=2||--workloads config1,large" \
  "fh4|-DCCU_FH_MIN_BLOCKS=4||--workloads config1,large"
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_full_err.log | tee gpurun_out/bench_full.json
tail -3 gpurun_out/bench_full_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_full_err.log | tee gpurun_out/bench_ref.json
