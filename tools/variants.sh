#!/bin/bash
# Build experiment variants of the library (headline size, single template only) into
# thrifty_b200/_lib/variants/<name>.so:
#   tools/variants.sh name1 "-DTHR_WL23=0" name2 "-D..." ...
# Select one at run time with THRIFTY_B200_LIB=<path>.
set -e
cd "$(dirname "$(readlink -f "$0")")/../thrifty_b200/csrc"
mkdir -p ../_lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( make -s -j5 OUTDIR=../_lib/variants/$name EXTRA="-DTHR_ONLY_N16384 $flags" 2>/dev/null >/dev/null \
    && cp ../_lib/variants/$name/libthrifty_b200.so ../_lib/variants/$name.so && echo "built $name" ) &
done
wait
ls -la ../_lib/variants/*.so
