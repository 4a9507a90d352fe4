#!/bin/bash
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py tests/test_host_renderer.py tests/test_gpu_multi.py -q -m gpu -x --timeout=600 2>&1 | tail -15 | tee gpurun_out/r2_pytest2.log
# BVH stage: shared-memory stack depth / leaf batch / walking warps
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities" \
  "stack12|-DCCU_Q_STACK=12||--workloads entities" \
  "stack16|-DCCU_Q_STACK=16||--workloads entities" \
  "stack24|-DCCU_Q_STACK=24||--workloads entities" \
  "leaf8||CCU_Q_LEAF_MIN=8|--workloads entities" \
  "leaf16||CCU_Q_LEAF_MIN=16|--workloads entities" \
  "bvhw26||CCU_Q_BVH_WARPS=26|--workloads entities" \
  "bvhw20||CCU_Q_BVH_WARPS=20|--workloads entities" \
  "norecs||CCU_NO_RECS=1|--workloads config1,indoor"
# ncu: full capture of the default kernel on config 1 (one 16-pass launch) and on the entity scene (2 passes)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 2 -c 1 -o gpurun_out/r02_queue_prof -f \
   python scripts/qbench.py --workloads config1 --reps 2 > gpurun_out/r02_queue_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 1 -c 1 -o gpurun_out/r02_queue_ent_prof -f \
   python scripts/qbench.py --workloads entities --reps 1 --passes 2 > gpurun_out/r02_queue_ent_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_first_hit -s 1 -c 1 -o gpurun_out/r02_first_hit_prof -f \
   python scripts/qbench.py --workloads config1 --reps 1 > gpurun_out/r02_first_hit_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
