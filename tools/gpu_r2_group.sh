#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_api.py -m gpu -q -x -p timeout --timeout 90 --timeout-method thread > gpurun_out/g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest.log
tail -15 gpurun_out/g_pytest.log | cut -c1-250
