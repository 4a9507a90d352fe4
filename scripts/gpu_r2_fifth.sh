#!/bin/bash
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -8 | tee gpurun_out/r2_pytest5.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "park|||--workloads entities,config1,large" \
  "nopark|-DCCU_BVH_PARK=0||--workloads entities" \
  "park_r4||CCU_Q_REFILL_MIN=4|--workloads entities" \
  "park_r12||CCU_Q_REFILL_MIN=12|--workloads entities" \
  "park_y12||CCU_YIELD_BELOW=12|--workloads entities" \
  "park_y26||CCU_YIELD_BELOW=26|--workloads entities" \
  "park_s12|-DCCU_Q_STACK=12||--workloads entities" \
  "park_s20|-DCCU_Q_STACK=20||--workloads entities"
timeout 900 python bench.py --workload large --spp-total 4096 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale_large_n1_err.log | tee gpurun_out/scale_large_n1.json
