#!/bin/bash
# one full ncu capture of the headline kernel (bench workload)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/${1:-r2}_full \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${1:-r2}_ncu_full.log 2>&1
tail -2 gpurun_out/${1:-r2}_ncu_full.log | cut -c1-200
ls -la gpurun_out/${1:-r2}_full.ncu-rep
