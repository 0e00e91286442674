#!/bin/bash
# Build a variant of the native library for an A/B run on the GPU box (never the product build):
#   tools/build_variant.sh NAME [-DMACRO ...]   ->  _exp/libdibs_b200_NAME.so
# Only dibs_abi.cu is recompiled with the extra macros; the per-DMAX Monte-Carlo objects are reused from dibs_b200/_build.
set -e
NAME=$1; shift
cd "$(dirname "$0")/.."
mkdir -p _exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c dibs_b200/csrc/dibs_abi.cu -o _exp/dibs_abi_$NAME.o
OBJS=$(ls dibs_b200/_build/mc_*.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static _exp/dibs_abi_$NAME.o $OBJS -o _exp/libdibs_b200_$NAME.so -ldl
echo _exp/libdibs_b200_$NAME.so
