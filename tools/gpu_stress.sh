#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tests/stress_parity.py ${1:-60} ${2:-1} > gpurun_out/stress.log 2>&1
grep -c "^ok" gpurun_out/stress.log; grep -A1 "^FAIL" gpurun_out/stress.log | cut -c1-700 | head -40; tail -1 gpurun_out/stress.log
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -k "stress or geometries" 2>&1 | tail -3
