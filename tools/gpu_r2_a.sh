#!/bin/bash
# round 2: full GPU test-suite and strict stress parity (plain 1e-4 bar, no widened carrier-offset tolerance)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p timeout --timeout 120 --timeout-method thread > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log | cut -c1-300
timeout 1200 python tests/stress_parity.py 200 99 > gpurun_out/a_stress.log 2>&1
grep -c "^ok" gpurun_out/a_stress.log; grep -A1 "^FAIL" gpurun_out/a_stress.log | cut -c1-500 | head -20; tail -1 gpurun_out/a_stress.log
timeout 600 python tests/stress_parity.py 40 3 big > gpurun_out/a_stress_big.log 2>&1
grep -c "^ok" gpurun_out/a_stress_big.log; grep -A1 "^FAIL" gpurun_out/a_stress_big.log | cut -c1-500 | head -10; tail -1 gpurun_out/a_stress_big.log
