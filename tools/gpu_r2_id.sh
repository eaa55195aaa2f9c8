#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_identify.py tests/test_gpu_group.py -m gpu -q -x -p timeout --timeout 90 --timeout-method thread > gpurun_out/i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/i_pytest.log
tail -25 gpurun_out/i_pytest.log | cut -c1-250
