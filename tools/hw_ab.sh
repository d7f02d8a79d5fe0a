#!/bin/bash
# half-width tail A/B on one box: the step's GEMM shapes at the row counts of 1 / 4 / 8-GPU ranks, FGB_GEMM_HW=0 vs 1, twice each
cd "$(dirname "$0")/.."
for rows in ${ROWS_LIST:-27280 13640 6820}; do
  for pass in 1 2; do
    for hw in 0 1; do
      for epi in 0 2; do echo -n "HW=$hw "; FGB_GEMM_HW=$hw ROWS=$rows python tools/gemm_sweep.py qkv,o,ffn1,ffn2 $epi 2>&1 | tail -1; done
    done
  done
done
