#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tests/stress_parity.py ${1:-40} ${2:-3} big > gpurun_out/stress_big.log 2>&1
grep -c "^ok" gpurun_out/stress_big.log; grep -c "detect2x" gpurun_out/stress_big.log; grep -A1 "^FAIL" gpurun_out/stress_big.log | cut -c1-600 | head -30; tail -1 gpurun_out/stress_big.log
