#!/bin/bash
# Builds a kernel-tuning variant of libries_b200.so into _ab/<name>.so (git-ignored; travels with gpurun).
# usage: tools/build_variant.sh <name> [-DEXP_THREADS=..] [-DEXP_PPT=..] ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=ries_b200/csrc
mkdir -p _ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-O2,-ffp-contract=off,-fvisibility=hidden -I $C "$@" -Xptxas=-v \
  -shared -Xcompiler -fvisibility=default -o _ab/$name.so $C/engine.cu $C/capi.cc $C/host/*.cc 2> _ab/$name.ptxas.log
grep -A2 "k_expand" _ab/$name.ptxas.log | grep -E "registers|spill" | head -3
