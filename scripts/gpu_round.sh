#!/bin/bash
# full regression + contract bench + launch list
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
timeout 1500 python -m pytest tests -q -m gpu -x --timeout=900 2>&1 | tail -6 | tee gpurun_out/r2_pytest19.log
timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench_err.log
tail -c 600 gpurun_out/r02c_bench_err.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_queue_bench_launches.csv \
   python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-other-workloads > gpurun_out/r02c_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_queue -s 2 -c 1 -o gpurun_out/r02d_queue_prof -f \
   python scripts/qbench.py --workloads config1 --reps 2 > gpurun_out/r02d_queue_ncu.log 2>&1
