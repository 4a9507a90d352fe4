#!/bin/bash
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -5 | tee gpurun_out/r2_pytest7.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "sel|||--workloads entities" \
  "sel_r4||CCU_Q_REFILL_MIN=4|--workloads entities" \
  "sel_r16||CCU_Q_REFILL_MIN=16|--workloads entities" \
  "sel_y28||CCU_YIELD_BELOW=28|--workloads entities" \
  "sel_y8||CCU_YIELD_BELOW=8|--workloads entities"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 1 -c 1 -o gpurun_out/r02_queue_ent_prof -f \
   python scripts/qbench.py --workloads entities --reps 1 --passes 2 > gpurun_out/r02_queue_ent_ncu.log 2>&1
