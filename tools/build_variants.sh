#!/bin/bash
# Builds A/B variants of the library into build/variants/<name>.so
# usage: tools/build_variants.sh name "-DDEN_KU=8 -DDEN_DOUBLE_BUFFER=0" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../torchain_b200/csrc"
mkdir -p ../../build/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  out=../../build/variants/$name
  mkdir -p $out
  for f in chain_api den_kernel num_kernel layout_kernel; do
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c -o $out/$f.o $f.cu &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/$name.so $out/*.o -cudart static
  echo built $name
done
