#!/bin/bash
# 8-GPU headline runs with every rank's per-kernel times (diagnostic for the scaling loss): default layout with the extras
# (layouts + parity: the record line), then the same without the stream-K tail
cd "$(dirname "$0")/.."
tag=${1:-r02_bench_n8_b}
nvidia-smi --query-gpu=index,power.limit,enforced.power.limit,power.max_limit,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/${tag}_smi.csv 2>&1
run() { out=$1; shift; env FGB_BENCH_RANK_KERNELS=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --warmup 3 ${EXTRA} > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; }
EXTRA="--steps 10" run ${tag}
EXTRA="--steps 6 --no-extras" run ${tag}_sk0 FGB_GEMM_SK=0
