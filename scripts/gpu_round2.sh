#!/bin/bash
# parity tests, smoke, bench (default + the other BASELINE workloads), ncu evidence for the default kernel
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_ours.json
tail -5 gpurun_out/bench_err.log
for w in indoor entities large; do
  timeout 900 python bench.py --steps 3 --warmup 3 --workload $w 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_$w.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render -s 3 -c 1 -f -o gpurun_out/queue_bench_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/queue_ncu.log 2>&1
tail -3 gpurun_out/queue_ncu.log
ls -la gpurun_out/ | tail -12
