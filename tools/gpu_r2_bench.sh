#!/bin/bash
# smoke + full bench line (1 GPU) + reference arm
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err || { echo BENCH FAILED; tail -5 gpurun_out/b_bench.err; }
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_bench.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'sustained',d.get('sustained',{}).get('value'))
print('cli',d.get('cli_e2e')); print('roofline',d['roofline']['frac'],d['roofline']['fp32']); print('cpu',d.get('cpu_baseline',{}).get('value'))
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err; cut -c1-400 gpurun_out/b_bench_ref.json
