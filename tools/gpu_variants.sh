#!/bin/bash
# usage: gpu_variants.sh name1 name2 ...   (libraries built by tools/variants.sh)
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
for v in "$@"; do
  THRIFTY_B200_LIB=$PWD/thrifty_b200/_lib/variants/$v.so timeout 300 python tests/variant_check.py >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
python - <<'PY'
import json
for l in open('gpurun_out/variants.jsonl'):
    d = json.loads(l)
    if 'parity_ok' in d:
        print(d['variant'].split('/')[-1], 'parity', d['parity_ok'], '' if d['parity_ok'] else d['detail'][:300])
    else:
        print('   %-40s %8.1f Gs/s' % (d['label'].split('/')[-1], d['msamples_per_s'] / 1e3))
PY
tail -3 gpurun_out/variants.err
