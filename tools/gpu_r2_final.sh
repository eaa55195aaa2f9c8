#!/bin/bash
# Round-2 evidence run (1 GPU): full GPU test-suite, strict stress parity, sweep, bench (ours + reference arm),
# ncu launch list + full capture of the headline kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/f_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p timeout --timeout 180 --timeout-method thread > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -4 gpurun_out/f_pytest.log | cut -c1-300
timeout 900 python tests/stress_parity.py 200 99 > gpurun_out/f_stress.log 2>&1
grep -c "^ok" gpurun_out/f_stress.log; grep -A1 "^FAIL" gpurun_out/f_stress.log | cut -c1-400 | head -8
timeout 600 python tests/stress_parity.py 100 5 fastdet > gpurun_out/f_stress_fd.log 2>&1; tail -1 gpurun_out/f_stress_fd.log
timeout 600 python tests/stress_parity.py 40 3 big > gpurun_out/f_stress_big.log 2>&1; tail -1 gpurun_out/f_stress_big.log
timeout 600 python tools/sweep.py > gpurun_out/f_sweep.jsonl 2> gpurun_out/f_sweep.err
timeout 600 python tools/sweep.py card >> gpurun_out/f_sweep.jsonl 2>> gpurun_out/f_sweep.err
cut -c1-170 gpurun_out/f_sweep.jsonl
timeout 900 python bench.py --steps 256 --warmup 8 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
cut -c1-600 gpurun_out/f_bench.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
cut -c1-300 gpurun_out/f_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv \
   python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 0 --no-cli > gpurun_out/f_ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/f_full \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 0 --no-cli > gpurun_out/f_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
