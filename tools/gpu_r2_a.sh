#!/bin/bash
# round 2, call A: GPU test-suite, strict stress parity (plain 1e-4 bar), short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
cut -c1-400 gpurun_out/a_bench.json
timeout 1500 python tests/stress_parity.py 200 99 > gpurun_out/a_stress.log 2>&1
grep -c "^ok" gpurun_out/a_stress.log; grep -A1 "^FAIL" gpurun_out/a_stress.log | cut -c1-500 | head -40; tail -1 gpurun_out/a_stress.log
