#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fastdet.py -m gpu -q 2>&1 | tail -3
for cfg in "8192" "4096" "32768" "16384 fastdet"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:detect -s 3 -c 1 -f -o gpurun_out/n_$tag python tools/sweep_one.py $cfg > gpurun_out/n_$tag.log 2>&1
  tail -1 gpurun_out/n_$tag.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
