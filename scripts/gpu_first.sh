#!/bin/bash
# first GPU contact: build checks + CUDA-vs-oracle parity
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0))"
make -C oracle CC=gcc 2>&1 | tail -2
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30
