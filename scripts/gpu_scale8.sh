#!/bin/bash
# scaling check on an 8-GPU box: bench at N = 1, 2, 4, 8 (weak scaling) + the 4K depth-11 world at N = 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale_err.log | tee gpurun_out/scale_n1.json | cut -c1-200
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/scale_err.log | tee gpurun_out/scale_n$n.json | cut -c1-200; fi
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --workload large 2>>gpurun_out/scale_err.log | tee gpurun_out/scale_large_n8.json | cut -c1-200
tail -3 gpurun_out/scale_err.log
