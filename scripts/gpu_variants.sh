#!/bin/bash
# usage: gpu_variants.sh "name|nvcc extra flags|env assignments|qbench args" ...   builds each variant next to the product library
# (in parallel) and times it with scripts/qbench.py.  Every variant of one workload must print the same md5.
mkdir -p gpurun_out /tmp/ccu_variants
pids=()
i=0
for spec in "$@"; do
  IFS='|' read -r name flags envs qargs <<< "$spec"
  if [ -n "$flags" ] || [ ! -f /tmp/ccu_variants/base.so ]; then
    out=/tmp/ccu_variants/$name.so
    [ -z "$flags" ] && out=/tmp/ccu_variants/base.so
    if [ ! -f "$out" ]; then
      ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
          -Xcompiler -fPIC -shared -cudart static --threads 2 $flags -o $out chunkyclplugin_b200/csrc/*.cu -ldl || echo "BUILD FAILED $name" ) &
      pids+=($!)
    fi
  fi
done
for p in "${pids[@]}"; do wait $p; done
for spec in "$@"; do
  IFS='|' read -r name flags envs qargs <<< "$spec"
  lib=/tmp/ccu_variants/$name.so
  [ -z "$flags" ] && lib=/tmp/ccu_variants/base.so
  [ -f "$lib" ] || continue
  env $envs CHUNKYCU_LIB=$lib timeout 120 python scripts/qbench.py --tag "$name" ${qargs:---workloads config1} 2>&1 | grep -v "^$" | tee -a gpurun_out/variants.jsonl
done
