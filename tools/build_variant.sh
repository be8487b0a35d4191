#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc -D flags...]  -> jlama_b200/variants/<name>.so (diagnostic builds)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
O=/tmp/variant_$name; mkdir -p $O jlama_b200/variants
for f in jl_runtime jl_gemv jl_gemm_tc jl_elementwise jl_attention jl_model jl_comm jl_mega; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c jlama_b200/csrc/$f.cu -o $O/$f.o &
done
wait
nvcc -shared -o jlama_b200/variants/$name.so $O/*.o -ldl -gencode arch=compute_100a,code=sm_100a
echo built jlama_b200/variants/$name.so
