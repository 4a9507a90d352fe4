#!/bin/bash
for l in 8 12 16 20; do for r in 4 8; do echo -n "entities bvh_warps=24 leaf_min=$l refill=$r: "; CCU_Q_REFILL_MIN=$r CCU_Q_LEAF_MIN=$l CCU_Q_BVH_WARPS=24 timeout 300 python scripts/run_render.py --scene entities --passes 4 --windows 2 --kernel 4 | grep "window 1"; done; done
