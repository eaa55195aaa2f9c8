#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python tools/sweep.py card > gpurun_out/io.jsonl 2> gpurun_out/io.err
cut -c1-420 gpurun_out/io.jsonl; tail -3 gpurun_out/io.err
