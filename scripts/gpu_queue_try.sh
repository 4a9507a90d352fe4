#!/bin/bash
make -C oracle CC=gcc >/dev/null
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for y in 16 20; do echo -n "terrain yield=$y: "; CCU_YIELD_BELOW=$y timeout 300 python scripts/run_render.py --passes 8 --windows 2 --kernel 4 | grep "window 1"; done
echo -n "indoor: "; timeout 300 python scripts/run_render.py --scene indoor --passes 8 --windows 2 --kernel 4 | grep "window 1"
echo -n "entities: "; timeout 300 python scripts/run_render.py --scene entities --passes 4 --windows 2 --kernel 4 | grep "window 1"
