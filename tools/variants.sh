#!/bin/bash
# Build experiment variants of the library (headline size only) into thrifty_b200/_lib/variants/<name>.so
#   tools/variants.sh name1 "-DTHR_WL23=0 -DTHR_ARGMAX1=0" name2 "-D..." ...
# Select one at run time with THRIFTY_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../thrifty_b200/csrc"
mkdir -p ../_lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
       -DTHR_ONLY_N16384 $flags -o ../_lib/variants/$name.so thrifty_b200.cu -lcudart 2>/dev/null &
done
wait
ls -la ../_lib/variants/
