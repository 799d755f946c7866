#!/bin/bash
# Builds libgraphwalk variants with different SimRank block sizes / samples-in-flight per thread
# into tools/variants/ (git-ignored) for A/B timing on the GPU box:  GW_LIB_OVERRIDE=tools/variants/libgw_b384_i2.so
set -e
cd "$(dirname "$0")/../graph_embedding_b200"
mkdir -p ../tools/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default"
for v in "$@"; do
  b=${v%%:*}; i=${v##*:}
  nvcc $FLAGS $EXTRA -DSR_BLOCK_THREADS=$b -DSR_ILP=$i -Xptxas -v -c csrc/simrank.cu -o ../tools/variants/simrank_b${b}_i${i}${TAG}.o 2>&1 | grep -A1 -E "k_simrank_logILi5" | grep -E "registers|spill" | tr '\n' ' '
  echo " <- block $b ilp $i"
  nvcc -shared -o ../tools/variants/libgw_b${b}_i${i}${TAG}.so build/graph.o build/alias.o build/walk.o build/walk_cn.o ../tools/variants/simrank_b${b}_i${i}${TAG}.o -gencode arch=compute_100a,code=sm_100a
done
