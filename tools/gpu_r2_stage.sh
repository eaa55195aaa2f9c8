#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -p timeout --timeout 60 --timeout-method thread > gpurun_out/s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s_pytest.log
tail -30 gpurun_out/s_pytest.log | cut -c1-250
