#!/bin/bash
# usage: gpu_sanitize.sh [block_len ...]   (default: all configurations of tools/sanitize_smoke.py)
: > gpurun_out/sanitizer.txt 2>/dev/null
mkdir -p gpurun_out
export THRIFTY_B200_MAX_GRID=2
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/sanitizer.txt
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_smoke.py "$@" > gpurun_out/san_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" gpurun_out/san_$tool.log | sort | uniq -c | head -12 >> gpurun_out/sanitizer.txt
  grep -E "^(16384|8192|4096|32768) " gpurun_out/san_$tool.log | wc -l >> gpurun_out/sanitizer.txt
done
cat gpurun_out/sanitizer.txt
