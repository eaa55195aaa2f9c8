#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ts_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/ts_pytest.log
tail -30 gpurun_out/ts_pytest.log
timeout 600 python tools/sweep.py > gpurun_out/ts_sweep.jsonl 2> gpurun_out/ts_sweep.err
cut -c1-200 gpurun_out/ts_sweep.jsonl; tail -3 gpurun_out/ts_sweep.err
