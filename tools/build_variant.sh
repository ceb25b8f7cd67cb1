#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc -D flags]  -> owlraytracing_b200/lib/<name>.so (kernel A/B experiments, tools/ab.py)
set -e
cd "$(dirname "$0")/../owlraytracing_b200/csrc"
name=$1; shift
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -diag-suppress 177 -shared --cudart shared \
  -Xlinker -rpath,/usr/local/cuda/lib64 -o ../lib/$name.so "$@" trueknn.cu dist.cu ingest.cpp -ldl
