#!/bin/bash
# round-2 first GPU trip: full GPU test-suite (incl. acceptance vs the reference kernel), smoke, quick kernel timings
mkdir -p gpurun_out
make -C oracle CC=gcc >/dev/null
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -q -m gpu -x --timeout=900 2>&1 | tail -25 > gpurun_out/r2_pytest.log
tail -25 gpurun_out/r2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python scripts/qbench.py --workloads config1,indoor,entities,large 2>&1 | tee gpurun_out/r2_qbench_first.json
