#!/bin/bash
# compare register budgets: rebuild with CCU_MIN_BLOCKS=N and time kernels 2/3
for nb in 2 3 4; do
  CCU_NVCC_EXTRA="-DCCU_MIN_BLOCKS=$nb" python chunkyclplugin_b200/build.py --force >/dev/null
  for k in 3 2; do
    echo -n "min_blocks=$nb kernel=$k: "; CCU_BLOCKS_PER_SM=$nb CCU_EXIT_IDLE=${EI:-16} python scripts/run_render.py --passes 8 --windows 2 --kernel $k | grep "window 1"
  done
done
python chunkyclplugin_b200/build.py --force >/dev/null
