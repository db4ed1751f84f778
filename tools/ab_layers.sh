#!/bin/bash
# A/B of library variants on the conv layer micro-benchmark, alternating in one process sequence (same box, same thermal state)
#   bash tools/ab_layers.sh "<lib A> <lib B> ..." "8 8 4096 8192" "16 16 2048 8192" ...
libs=$1; shift
for rep in 1 2; do
  for c in "$@"; do
    for l in $libs; do
      if [ "$l" = main ]; then unset POPCORN_B200_LIB; else export POPCORN_B200_LIB=$PWD/tools/probe/$l; fi
      echo -n "[$l] "; KB_ONLY=tc KB_ITERS=20 python tools/conv_layer_bench.py $c | tail -1
    done
  done
done
