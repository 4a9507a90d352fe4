#!/bin/bash
make -C oracle CC=gcc >/dev/null
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "queue and (entities or mixed)" 2>&1 | tail -3
for l in 2 4 8 12 16; do for b in 23 25; do echo -n "entities leaf_min=$l bvh_warps=$b: "; CCU_Q_LEAF_MIN=$l CCU_Q_BVH_WARPS=$b timeout 300 python scripts/run_render.py --scene entities --passes 4 --windows 2 --kernel 4 | grep "window 1"; done; done
