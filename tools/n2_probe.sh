#!/bin/bash
# 2-GPU pure-Ulysses runs at 13 640 tokens (per-rank GEMM shapes = an SP4 rank of the headline): which kernels lose time under the exchange
cd "$(dirname "$0")/.."
run() { env FGB_BENCH_SHAPE=352x1280x121 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 4 --warmup 3 --layout sp --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', round(d['ms_per_step'],2),'ms', d['clocks']['sm_mhz'],'MHz', d['clocks']['power_w_max'],'W', json.dumps(d['kernel_ms_per_step']))"; }
run FGB_GEMM_SK=0
run FGB_GEMM_SK=1
run FGB_GEMM_SK=0 FGB_SP_FUSED=0
run FGB_GEMM_SK=0
run FGB_GEMM_SK=1
