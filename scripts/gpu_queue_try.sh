#!/bin/bash
for cfg in "26 20" "26 22" "30 22" "30 24" "32 24" "32 26"; do set -- $cfg
CCU_NVCC_EXTRA="-DCCU_Q_WARPS=$1" python chunkyclplugin_b200/build.py --force >/dev/null
echo -n "warps=$1 march_warps=$2: "; CCU_Q_MARCH_WARPS=$2 timeout 120 python scripts/run_render.py --passes 8 --windows 2 --kernel 4 | grep "window 1"
done
python chunkyclplugin_b200/build.py --force >/dev/null
