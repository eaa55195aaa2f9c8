#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c2_pytest.log
tail -5 gpurun_out/c2_pytest.log
timeout 300 python tools/variant_check.py > gpurun_out/c2_variants.jsonl 2> gpurun_out/c2_variants.err
cut -c1-330 gpurun_out/c2_variants.jsonl; tail -3 gpurun_out/c2_variants.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/c2_full \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/c2_ncu.log 2>&1
