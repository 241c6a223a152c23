#!/bin/bash
# build a variant of the library with extra -D flags into gpurun_out/variants/<name>.so  (development tool)
# usage: tools/build_variant.sh name -DGG_BWD_KUNROLL=2 ...
set -e
name=$1; shift
mkdir -p variants
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3,-fvisibility=hidden --expt-relaxed-constexpr -shared "$@" \
  -o variants/$name.so gaussigan_b200/csrc/*.cu -Xptxas -v 2>&1 | grep -E "error|bwd_mask_fast_kernelILi1ELb0ELi8" -A2 | grep -E "error|Used|spill"
