#!/bin/bash
# usage: gpu_ab2.sh "<libA> <block_len> [mode]" "<libB> <block_len> [mode]" ...   (lib = variant name or "default")
# Times tools/sweep_one.py for each entry, twice, interleaved; every run under a short timeout.
entries=("$@")
for rep in 1 2; do
  for e in "${entries[@]}"; do
    read -r lib n mode <<< "$e"
    if [ "$lib" = default ]; then unset THRIFTY_B200_LIB; else export THRIFTY_B200_LIB=$PWD/thrifty_b200/_lib/variants/$lib.so; fi
    printf "%-10s N=%-6s %-8s " "$lib" "$n" "$mode"
    timeout 60 python tools/sweep_one.py $n $mode 2>&1 | grep -o '"msamples_per_s": [0-9.]*' || echo "failed or hung"
  done
done
