#!/bin/bash
for m in 64 26 24 22 20; do for y in 8 16; do echo -n "terrain march_warps=$m yield=$y: "; CCU_YIELD_BELOW=$y CCU_Q_MARCH_WARPS=$m timeout 300 python scripts/run_render.py --passes 8 --windows 2 --kernel 4 | grep "window 1"; done; done
