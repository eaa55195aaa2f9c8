#!/bin/bash
# 2 x 16384 kernel (block_len 32768): parity tests of both stage-A modes under short timeouts, then timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_api.py -m gpu -q -x -p timeout --timeout 90 --timeout-method thread -k "n32768" > gpurun_out/x_2x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/x_2x_pytest.log
tail -6 gpurun_out/x_2x_pytest.log | cut -c1-400
timeout 120 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tools")
import numpy as np, sweep
ex = np.load("tests/golden/template_example.npy")
sweep.run(32768, ex, 4920, 2048, 1.0, label="N=32768 zoom")
sweep.run(32768, ex, 4920, 2048, 1.0, window=(7, 300), label="N=32768 full FFT#1 (2x kernel)")
PY
