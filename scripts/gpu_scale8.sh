#!/bin/bash
# 8-GPU box: group tests, weak-scaling config 1 at N=8, strong-scaling config 5 (4096 spp) at N=8, and the N=1 lines on the same box
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --timeout=600 2>&1 | tail -5 | tee gpurun_out/r2c_pytest_multi8.log
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
run 8 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02c_scale_n8_err.log > gpurun_out/r02c_scale_n8.json
tail -3 gpurun_out/r02c_scale_n8_err.log | cut -c1-300
run 8 --workload large --spp-total 4096 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02c_scale_large_n8_err.log > gpurun_out/r02c_scale_large_n8.json
tail -3 gpurun_out/r02c_scale_large_n8_err.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-workloads 2>gpurun_out/r02c_scale_n1_err.log > gpurun_out/r02c_scale_n1.json
timeout 600 python bench.py --workload large --spp-total 4096 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02c_scale_large_n1_err.log > gpurun_out/r02c_scale_large_n1.json
for f in gpurun_out/r02c_scale_n8.json gpurun_out/r02c_scale_large_n8.json gpurun_out/r02c_scale_n1.json gpurun_out/r02c_scale_large_n1.json; do python -c "
import json,sys
l=json.loads(open('$f').read().strip().split('\n')[-1]); print('$f', l['n_gpus'], l['value']/1e9, l['e2e']['value']/1e9, l['ms_per_step'])"; done
