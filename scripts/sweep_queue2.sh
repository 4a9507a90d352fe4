#!/bin/bash
# usage: sweep_queue2.sh "<nvcc extra flags variant>;..." [env assignments]
IFS=';' read -ra VARS <<< "$1"
for v in "${VARS[@]}"; do
  CCU_NVCC_EXTRA="$v" python chunkyclplugin_b200/build.py --force >/dev/null || { echo "build failed: $v"; continue; }
  for r in ${RS:-8}; do for y in ${YS:-24}; do
  echo -n "[$v] refill=$r yield=$y ${EXTRA}: "; CCU_Q_REFILL_MIN=$r CCU_YIELD_BELOW=$y timeout 120 python scripts/run_render.py --passes 8 --windows 2 --kernel 4 $EXTRA | grep "window 1"
  done; done
done
python chunkyclplugin_b200/build.py --force >/dev/null
