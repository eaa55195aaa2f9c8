#!/bin/bash
mkdir -p gpurun_out
for seed in 7 123 2026; do
  timeout 900 python tests/stress_parity.py 200 $seed > gpurun_out/x_stress_$seed.log 2>&1
  echo "seed $seed: $(grep -c '^ok' gpurun_out/x_stress_$seed.log) ok; $(tail -1 gpurun_out/x_stress_$seed.log)"
  grep -A1 "^FAIL" gpurun_out/x_stress_$seed.log | cut -c1-420 | head -12
done
timeout 600 python tests/stress_parity.py 60 11 big > gpurun_out/x_stress_big.log 2>&1; tail -1 gpurun_out/x_stress_big.log; grep -A1 "^FAIL" gpurun_out/x_stress_big.log | cut -c1-420 | head -8
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -k "pipelined_chunks" 2>&1 | tail -2
