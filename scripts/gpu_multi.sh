#!/bin/bash
# multi-GPU check: NCCL test + scaling bench lines for N = 1, 2 (and more if present)
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
NG=$(nvidia-smi -L | wc -l)
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_multi_err.log | tee gpurun_out/bench_n1.json
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_multi_err.log | tee gpurun_out/bench_n$n.json
  fi
done
tail -5 gpurun_out/bench_multi_err.log
