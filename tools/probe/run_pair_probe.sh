#!/bin/bash
# Round-2 first GPU call:   gpurun --timeout 300 -- 'bash tools/probe/run_pair_probe.sh > gpurun_out/pair_probe.txt 2>&1'
# Every variant runs in its own process under its own timeout (a faulting descriptor poisons the CUDA context; a wrong one
# must never hang the box: the probe's barrier waits trap after a bounded spin).
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I popcorn_b200/csrc -o tools/probe/pair_probe tools/probe/pair_probe.cu -lcuda || exit 1
for variant in 0 1 2 3; do
  for swap in 0 1; do
    echo "=== variant $variant swap_lbo_sbo $swap ==="
    timeout 20 tools/probe/pair_probe $swap $variant
    echo "exit code $?"
  done
done
