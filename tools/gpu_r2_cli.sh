#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_api.py -m gpu -q -x -p timeout --timeout 120 --timeout-method thread -k "cli or card or stream" > gpurun_out/c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -8 gpurun_out/c_pytest.log | cut -c1-300
df -h /dev/shm | tail -1
timeout 300 python tools/cli_profile.py 16384 2>&1 | head -40
