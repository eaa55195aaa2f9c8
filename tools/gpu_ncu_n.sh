#!/bin/bash
# ncu --set full of one launch of the kernel for block length $1 (tools/sweep_one.py workload); report name suffix $2
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/n$1_$2 \
   python tools/sweep_one.py $1 > gpurun_out/n$1_$2.log 2>&1
tail -2 gpurun_out/n$1_$2.log | cut -c1-200
