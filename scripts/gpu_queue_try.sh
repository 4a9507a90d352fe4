#!/bin/bash
for b in 20 22 23 24 25; do for m in 22 28; do echo -n "entities bvh_warps=$b march_warps=$m: "; CCU_Q_BVH_WARPS=$b CCU_Q_MARCH_WARPS=$m timeout 300 python scripts/run_render.py --scene entities --passes 4 --windows 2 --kernel 4 | grep "window 1"; done; done
