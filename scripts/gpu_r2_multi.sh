#!/bin/bash
# 2-GPU box: group tests, bench at N=1 and N=2
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_features.py tests/test_host_renderer.py -q -m gpu -x --timeout=600 2>&1 | tail -15 | tee gpurun_out/r2_pytest_multi.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-other-workloads 2>gpurun_out/bench_n1_err.log | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_n1_err.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2_err.log | tee gpurun_out/bench_n2.json
tail -5 gpurun_out/bench_n2_err.log
rm -f gpurun_out/variants.jsonl
bash scripts/gpu_variants.sh \
  "base|||--workloads config1,indoor,large" \
  "nosky||CCU_NO_SKY_SMEM=1|--workloads config1,indoor"
