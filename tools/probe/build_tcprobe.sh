#!/bin/bash
# Probe builds of the library (conv_tc.cu with -DPC_TC_PROBE=1 and extra defines): tools/probe/libpc_tcprobe<tag>.so
#   bash tools/probe/build_tcprobe.sh <tag> [-DPC_TC_EXP=1 ...]      (run `make -C popcorn_b200/csrc` first: the other objects are reused)
set -e
cd "$(dirname "$0")/../../popcorn_b200/csrc"
tag=$1; shift
NV=/usr/local/cuda/bin/nvcc
$NV -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DPC_TC_PROBE=1 "$@" -c conv_tc.cu -o /tmp/conv_tc_probe$tag.o
$NV -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/probe/libpc_tcprobe$tag.so api.o conv.o /tmp/conv_tc_probe$tag.o head.o head_tc.o head_bwd.o region.o ingest.o unet_bwd.o
