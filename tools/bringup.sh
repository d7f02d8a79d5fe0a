#!/bin/bash
# GPU bring-up sweep: every case in its own process with a timeout so a trap/hang cannot take the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K=tools/kcheck
run() { echo "== $*"; timeout 60 $K "$@" 2>&1 | tail -5; echo "exit=$?"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for epi in 0 1 2 3; do run gemm 128 256 64 $epi; done
run gemm 128 256 256 0
run gemm 256 512 512 0
run gemm 200 192 192 0
run gemm 2 768 256 0
run gemm 1000 3072 3072 2
run gemm 27280 3072 3072 0 10
run gemm 27280 9216 3072 0 10 0
run gemm 27280 14336 3072 1 10 0
run gemm 27280 3072 14336 2 10 0
run attn 256 128 1
run attn 256 256 1
run attn 128 512 2
run attn 300 200 2
run attn 1000 1000 3
run attn 2048 2048 2 5
run attn 27280 512 24 5 0
run attn 27280 27280 24 3 0
