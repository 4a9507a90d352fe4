#!/bin/bash
# rebuild the queue kernel with different pool shapes and time it.  usage: sweep_queue.sh "W:R ..." "yield ..." "refill ..." [extra run_render args]
CFGS=${1:-"32:48"}
YS=${2:-"24"}
RS=${3:-"1 8"}
EXTRA=${4:-""}
for cfg in $CFGS; do
  W=${cfg%%:*}; R=${cfg##*:}
  CCU_NVCC_EXTRA="-DCCU_Q_WARPS=$W -DCCU_Q_ROWS=$R" python chunkyclplugin_b200/build.py --force >/dev/null || { echo "build failed $cfg"; continue; }
  for y in $YS; do for r in $RS; do
    echo -n "warps=$W rows=$R yield_below=$y refill_min=$r $EXTRA: "; CCU_Q_REFILL_MIN=$r CCU_YIELD_BELOW=$y timeout 120 python scripts/run_render.py --passes 8 --windows 2 --kernel 4 $EXTRA | grep "window 1"
  done; done
done
python chunkyclplugin_b200/build.py --force >/dev/null
