#!/bin/bash
# quick check: headline golden parity + short bench (and optional extra pytest -k expression in $1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x ${1:+-k "$1"} > gpurun_out/q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'clk',d['clocks'])
PY
