#!/bin/bash
for w in 8 12 16 20 24 28; do echo -n "kernel3 wait=$w: "; CCU_WAIT_LANES=$w python scripts/run_render.py --passes 8 --windows 2 --kernel 3 | grep "window 1"; done
for e in 8 16 24; do echo -n "kernel2 exit_idle=$e: "; CCU_EXIT_IDLE=$e python scripts/run_render.py --passes 8 --windows 2 --kernel 2 | grep "window 1"; done
