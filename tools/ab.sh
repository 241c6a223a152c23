#!/bin/bash
# A/B of library variants built by tools/build_variant.sh (development tool): per-kernel times and the headline step chain
# usage (on the GPU box): tools/ab.sh name1 name2 ...   -> gpurun_out/ab_<name>.txt
mkdir -p gpurun_out
for v in "$@"; do
  ( GAUSSIGAN_LIB=$PWD/variants/$v.so python tools/ktime.py 2>&1 | head -12
    TAG=$v GAUSSIGAN_LIB=$PWD/variants/$v.so python tools/time_step.py 2>&1 | tail -2 ) > gpurun_out/ab_$v.txt
  echo "== $v"; cat gpurun_out/ab_$v.txt
done
