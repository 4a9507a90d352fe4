#!/bin/bash
for b in 0 4 8 16 32; do echo -n "march_bias=$b: "; CCU_Q_MARCH_BIAS=$b timeout 60 python scripts/run_render.py --passes 8 --windows 2 --kernel 4 | grep "window 1"; done
