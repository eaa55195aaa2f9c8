#!/bin/bash
# quick check: smoke under a short timeout (a dead-locked kernel must not eat the budget), headline golden parity,
# short bench (and optional extra pytest -k expression in $1).  Every step dies after 60-150 s.
mkdir -p gpurun_out
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "SMOKE FAILED/HUNG"; exit 1; }
timeout 120 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err || { echo "BENCH FAILED/HUNG"; tail -3 gpurun_out/q_bench.err; exit 1; }
python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'clk',d['clocks'])
PY
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x -p timeout --timeout 60 --timeout-method thread ${1:+-k "$1"} > gpurun_out/q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log | cut -c1-300
