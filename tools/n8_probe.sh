#!/bin/bash
# 8-GPU headline run with every rank's per-kernel times: default layout with the extras (layouts + parity) = the record line
cd "$(dirname "$0")/.."
tag=${1:-r02_bench_n8}
nvidia-smi --query-gpu=index,power.limit,enforced.power.limit,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/${tag}_smi.csv 2>&1
env FGB_BENCH_RANK_KERNELS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --warmup 3 --steps 10 > gpurun_out/$tag.json 2> gpurun_out/$tag.err; echo "$tag rc=$?"
