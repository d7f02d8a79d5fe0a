#!/bin/bash
cd "$(dirname "$0")/.."
K=tools/kcheck
run() { timeout 90 $K "$@" 2>&1 | grep -E "rel_l2|TFLOP|error|Error"; }
run gemm 1000 3072 3072 2
run gemm 200 192 192 0
run gemm 27280 3072 3072 0 20 0
run gemm 27280 9216 3072 0 20 0
run gemm 27280 14336 3072 1 20 0
run gemm 27280 3072 14336 2 20 0
for e in ${EMUS:-0 2 3 4}; do
  echo "== EMU=$e"; FGB_ATTN_EMU=$e run attn 1000 1000 3
  FGB_ATTN_EMU=$e run attn 27280 27280 24 5 0
done
for a in "256 128 1" "300 200 2" "128 512 2" "2048 2048 2"; do run attn $a; done
run attn 27280 512 24 5 0
