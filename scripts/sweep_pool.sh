#!/bin/bash
# tuning sweep of the path-pool kernel's knobs
for r in 1 4 8; do for e in 2 4 8 16 24; do
  echo -n "refill_min=$r exit_idle=$e: "; CCU_REFILL_MIN=$r CCU_EXIT_IDLE=$e python scripts/run_render.py --passes 8 --windows 2 | grep "window 1"
done; done
echo -n "wavefront(lane-bound) wait=24: "; python scripts/run_render.py --passes 8 --windows 2 --kernel 3 | grep "window 1"
echo -n "megakernel: "; python scripts/run_render.py --passes 8 --windows 2 --kernel 1 | grep "window 1"
