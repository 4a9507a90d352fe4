#!/bin/bash
make -C oracle CC=gcc >/dev/null
RS="8" YS="16 20 24" bash scripts/sweep_queue2.sh "-DCCU_Q_PARTNER"
CCU_NVCC_EXTRA="-DCCU_Q_PARTNER" python chunkyclplugin_b200/build.py --force >/dev/null
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "queue" 2>&1 | tail -3
python chunkyclplugin_b200/build.py --force >/dev/null
