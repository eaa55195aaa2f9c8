#!/bin/bash
# full GPU test-suite; log merged back as gpurun_out/pytest_gpu.log
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-220
