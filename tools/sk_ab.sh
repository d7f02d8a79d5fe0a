#!/bin/bash
# stream-K tail A/B on one box: the step's GEMM shapes at the row counts of 1 / 4 / 8-GPU ranks, unsplit vs split, each twice
cd "$(dirname "$0")/.."
for rows in ${ROWS_LIST:-27280 13640 6820 3410}; do
  for pass in 1 2; do
    for sk in 0 1; do
      for epi in 0 2; do ROWS=$rows SK=$sk python tools/gemm_sweep.py qkv,o,ffn1,ffn2 $epi 2>&1 | tail -1; done
    done
  done
done
