#!/bin/bash
mkdir -p gpurun_out /tmp/ccu_variants
make -C oracle CC=gcc >/dev/null
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -q -m gpu -x --timeout=600 2>&1 | tail -4 | tee gpurun_out/r2_pytest16.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,entities,indoor" \
  "unroll2|-DCCU_MARCH_UNROLL=2||--workloads config1,entities,indoor" \
  "unroll2_r6|-DCCU_MARCH_UNROLL=2|CCU_Q_REFILL_MIN=6|--workloads config1" \
  "unroll2_r10|-DCCU_MARCH_UNROLL=2|CCU_Q_REFILL_MIN=10|--workloads config1" \
  "unroll3|-DCCU_MARCH_UNROLL=3||--workloads config1" \
  "ent_r4||CCU_Q_REFILL_MIN=4|--workloads entities" \
  "ent_r12||CCU_Q_REFILL_MIN=12|--workloads entities" \
  "ent_r16||CCU_Q_REFILL_MIN=16|--workloads entities" \
  "ent_y8||CCU_YIELD_BELOW=8|--workloads entities" \
  "ent_y28||CCU_YIELD_BELOW=28|--workloads entities" \
  "ent_stack8|-DCCU_Q_STACK=8||--workloads entities" \
  "ent_sticky8||CCU_Q_STICKY=8|--workloads entities" \
  "ent_sticky24||CCU_Q_STICKY=24|--workloads entities"
CCU_NVCC_EXTRA="-DCCU_Q_STATS" python chunkyclplugin_b200/build.py --force > /dev/null
python scripts/qbench.py --workloads config1,entities --reps 1 --passes 4 2> gpurun_out/r02c_qstats.txt | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 1 -c 1 -o gpurun_out/r02c_queue_ent_prof -f \
   env CHUNKYCU_LIB=/tmp/ccu_variants/base.so python scripts/qbench.py --workloads entities --reps 1 --passes 2 > gpurun_out/r02c_queue_ent_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 2 -c 1 -o gpurun_out/r02c_queue_prof -f \
   env CHUNKYCU_LIB=/tmp/ccu_variants/base.so python scripts/qbench.py --workloads config1 --reps 2 > gpurun_out/r02c_queue_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
