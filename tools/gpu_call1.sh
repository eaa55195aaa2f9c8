#!/bin/bash
# GPU call: full gpu test-suite on the default build, variant parity + timing, ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
for v in base wl wlam rho all; do
  THRIFTY_B200_LIB=$PWD/thrifty_b200/_lib/variants/$v.so timeout 300 python tools/variant_check.py >> gpurun_out/c1_variants.jsonl 2>> gpurun_out/c1_variants.err
done
cat gpurun_out/c1_variants.jsonl | cut -c1-400
timeout 600 python bench.py --steps 256 --warmup 8 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
cat gpurun_out/c1_bench.json | cut -c1-1500
timeout 600 ncu --set full --import-source on --clock-control none -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/c1_full \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/c1_ncu.log 2>&1
ls -la gpurun_out
