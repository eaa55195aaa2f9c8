#!/bin/bash
# multi-GPU evidence run (gpurun --gpus N): H2D ceiling of the node, group tests, bench at 2..N GPUs
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/m_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/m_lscpu.txt 2>&1
( cd tools/microbench && timeout 120 ./h2d_concurrent 0.6 128 ) > gpurun_out/m_h2d_concurrent.jsonl 2> gpurun_out/m_h2d.err
cat gpurun_out/m_h2d_concurrent.jsonl | cut -c1-220
cp gpurun_out/m_h2d_concurrent.jsonl profiles/r02_h2d_concurrent.jsonl
timeout 300 python -m pytest tests/test_gpu_group.py -m gpu -q -x -p timeout --timeout 90 --timeout-method thread > gpurun_out/m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_pytest.log
tail -3 gpurun_out/m_pytest.log | cut -c1-200
for n in $NG 4 2; do
  [ $n -le $NG ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $n --steps 64 --warmup 8 > gpurun_out/m_bench_$n.json 2> gpurun_out/m_bench_$n.err || { echo "bench $n failed"; tail -5 gpurun_out/m_bench_$n.err; }
  python - $n <<'PY'
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/m_bench_%s.json'%sys.argv[1]) if l.startswith('{')][-1])
    e=d['e2e']
    print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(e['value']),'h2d GB/s',round(e['h2d_gbs'],1),'ceiling',e['h2d_ceiling_gbs'],'frac',e['frac_of_h2d_ceiling'],'stripe_parity',d.get('stripe_parity'),'group==single',e.get('group_records_equal_single_gpu'),'numa',e.get('numa_nodes'))
except Exception as ex: print('parse failed',ex)
PY
done
