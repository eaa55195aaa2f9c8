#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python tools/sweep.py card > gpurun_out/c15_io.jsonl 2> gpurun_out/c15_io.err
cat gpurun_out/c15_io.jsonl | cut -c1-500; tail -3 gpurun_out/c15_io.err
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e'])"
