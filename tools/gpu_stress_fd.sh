#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tests/stress_parity.py ${1:-60} ${2:-1} fastdet > gpurun_out/stress_fd.log 2>&1
grep -c "^ok" gpurun_out/stress_fd.log; grep -A1 "^FAIL" gpurun_out/stress_fd.log | cut -c1-700 | head -40; tail -1 gpurun_out/stress_fd.log
