#!/bin/bash
make -C oracle CC=gcc >/dev/null
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "queue" 2>&1 | tail -3
for y in 16 24; do for r in 6 8; do echo -n "default build yield=$y refill=$r: "; CCU_REFILL_MIN=$r CCU_YIELD_BELOW=$y timeout 60 python scripts/run_render.py --passes 8 --windows 2 --kernel 4 | grep "window 1"; done; done
