#!/bin/bash
make -C oracle CC=gcc >/dev/null
timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo -n "benchmark scene (reference's Greenfield fixture) 1080p kernel4: "; timeout 300 python scripts/run_render.py --scene benchmark --passes 8 --windows 2 --kernel 4 | grep "window 1"
echo -n "benchmark scene kernel3: "; timeout 300 python scripts/run_render.py --scene benchmark --passes 8 --windows 2 --kernel 3 | grep "window 1"
echo -n "benchmark scene kernel1: "; timeout 300 python scripts/run_render.py --scene benchmark --passes 8 --windows 2 --kernel 1 | grep "window 1"
