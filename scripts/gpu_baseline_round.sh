#!/bin/bash
# re-establish GPU evidence: parity tests, smoke, bench (both arms), ncu launch list of the bench command, one full capture of the render kernel
mkdir -p gpurun_out
nproc; lscpu | grep "Model name"
make -C oracle CC=gcc >/dev/null
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_ours.json
tail -5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/wave_prof python scripts/run_render.py --passes 4 --windows 2 > gpurun_out/wave_ncu.log 2>&1
tail -3 gpurun_out/wave_ncu.log
ls -la gpurun_out/
