#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus $N --steps 256 --warmup 8 > gpurun_out/m_bench_$N.json 2> gpurun_out/m_bench_$N.err
cut -c1-400 gpurun_out/m_bench_$N.json; tail -2 gpurun_out/m_bench_$N.err | cut -c1-300
