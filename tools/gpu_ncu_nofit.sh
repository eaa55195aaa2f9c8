#!/bin/bash
mkdir -p gpurun_out
export THRIFTY_B200_LIB=$PWD/thrifty_b200/_lib/variants/nofit_all.so
timeout 300 ncu --set full --import-source on --clock-control none -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/n4096_nofit python tools/sweep_one.py 4096 > gpurun_out/n4096_nofit.log 2>&1
tail -1 gpurun_out/n4096_nofit.log | cut -c1-200
