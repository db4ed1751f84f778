#!/bin/bash
# A/B builds of the library: conv_tc.cu / head_tc.cu recompiled with extra defines -> tools/probe/libpc_<tag>.so
#   bash tools/probe/build_variant.sh <tag> [-DPC_TC_ST16=0 ...]      (run `make -C popcorn_b200/csrc` first: the other objects are reused)
set -e
cd "$(dirname "$0")/../../popcorn_b200/csrc"
tag=$1; shift
NV=/usr/local/cuda/bin/nvcc
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
$NV $FL "$@" -c conv_tc.cu -o /tmp/conv_tc_$tag.o
$NV $FL "$@" -c head_tc.cu -o /tmp/head_tc_$tag.o
$NV -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/probe/libpc_$tag.so api.o conv.o /tmp/conv_tc_$tag.o head.o /tmp/head_tc_$tag.o head_bwd.o region.o ingest.o unet_bwd.o
