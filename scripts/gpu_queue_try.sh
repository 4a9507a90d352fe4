#!/bin/bash
for cl in 4 6 8; do echo -n "large 4K cell_level=$cl: "; CCU_CELL_LEVEL=$cl timeout 600 python scripts/run_render.py --scene large --width 3840 --height 2160 --passes 4 --windows 2 --kernel 4 | grep "window 1"; done
