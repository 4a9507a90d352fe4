#!/bin/bash
# tuning sweep of the wavefront kernel's knobs
for b in 1 2 3; do for w in 1 4 8 12 16 24; do
  echo -n "blocks_per_sm=$b wait=$w: "; CCU_BLOCKS_PER_SM=$b CCU_WAIT_LANES=$w python scripts/run_render.py --passes 8 --windows 2 | grep "window 1"
done; done
echo -n "megakernel: "; python scripts/run_render.py --passes 8 --windows 2 --kernel 1 | grep "window 1"
