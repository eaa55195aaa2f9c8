#!/bin/bash
# 2-GPU check: bench under torchrun (ours + reference arm), plus the GPU test-suite on GPU 0
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 256 --warmup 8 > gpurun_out/m_bench_$N.json 2> gpurun_out/m_bench_$N.err
cut -c1-700 gpurun_out/m_bench_$N.json; tail -3 gpurun_out/m_bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/m_bench_ref_$N.json 2> gpurun_out/m_bench_ref_$N.err
cut -c1-300 gpurun_out/m_bench_ref_$N.json
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_pytest.log
tail -3 gpurun_out/m_pytest.log
