#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python tools/sweep_one.py 16384 fastdet 2>&1 | cut -c1-200
timeout 600 python tools/sweep.py card > gpurun_out/c14_io.jsonl 2> gpurun_out/c14_io.err
cat gpurun_out/c14_io.jsonl | cut -c1-500; tail -3 gpurun_out/c14_io.err
