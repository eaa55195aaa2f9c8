#!/bin/bash
# usage: gpu_ab_n.sh <variant name> <block_len> [pytest -k expression]
# A/B of a full-library variant (thrifty_b200/_lib/variants/<name>.so) against the default build at one block length.
# The variant is timed first under a short timeout: a dead-locked experiment must not eat the GPU budget.
mkdir -p gpurun_out
V=$PWD/thrifty_b200/_lib/variants/$1.so
N=$2
K=${3:-$N}
echo $1; THRIFTY_B200_LIB=$V timeout 60 python tools/sweep_one.py $N 2>&1 | grep -o '"msamples_per_s": [0-9.]*' || { echo "variant failed or hung"; exit 1; }
THRIFTY_B200_LIB=$V timeout 200 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -5
for i in 1 2; do
  echo default; timeout 60 python tools/sweep_one.py $N 2>&1 | grep -o '"msamples_per_s": [0-9.]*'
  echo $1; THRIFTY_B200_LIB=$V timeout 60 python tools/sweep_one.py $N 2>&1 | grep -o '"msamples_per_s": [0-9.]*'
  echo "$1 fastdet"; THRIFTY_B200_LIB=$V timeout 60 python tools/sweep_one.py $N fastdet 2>&1 | grep -o '"msamples_per_s": [0-9.]*'
  echo "default fastdet"; timeout 60 python tools/sweep_one.py $N fastdet 2>&1 | grep -o '"msamples_per_s": [0-9.]*'
done
