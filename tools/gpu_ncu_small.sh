#!/bin/bash
# ncu --set full of one launch of the N=8192 and N=4096 kernels (tools/sweep_one.py workloads)
mkdir -p gpurun_out
for n in 8192 4096; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/small_$n \
     python tools/sweep_one.py $n > gpurun_out/small_$n.log 2>&1
  tail -2 gpurun_out/small_$n.log | cut -c1-200
done
ls -la gpurun_out/small_*
