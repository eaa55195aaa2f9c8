#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "32768 or golden" > gpurun_out/c6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c6_pytest.log
tail -30 gpurun_out/c6_pytest.log | cut -c1-250
timeout 300 python - > gpurun_out/c6_sweep.jsonl 2> gpurun_out/c6_sweep.err <<'PY'
import sys, os, numpy as np
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import sweep
example = np.load('tests/golden/template_example.npy')
sweep.run(32768, example, 4920, 2048, 1.0, steps=32, label="cfg3 N=32768 (2x16384 kernel)")
PY
cut -c1-250 gpurun_out/c6_sweep.jsonl; tail -3 gpurun_out/c6_sweep.err
