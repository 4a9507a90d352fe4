#!/bin/bash
# usual GPU round trip: parity tests, smoke, bench (ours + reference arm)
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps ${STEPS:-5} --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_ours.json
tail -5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json
