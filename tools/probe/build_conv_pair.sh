#!/bin/bash
# Builds the round-2 candidate kernel into its own library (never linked into libpopcorn_b200.so).
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr \
     -I popcorn_b200/csrc -shared -o popcorn_b200/libpopcorn_b200_probe.so tools/probe/conv_pair.cu tools/probe/conv_ss.cu tools/probe/c4_kernels.cu tools/probe/dda_c4.cu -lcuda "$@"
