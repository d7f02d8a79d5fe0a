#!/bin/bash
# supertile sweep of the 2-CTA GEMM (one process per setting: the library reads the environment once)
cd "$(dirname "$0")/.."
python tools/gemm_sweep.py qkv,o,ffn1,ffn2
for bn in 12 6 4 3; do for gm in 2 4 6 9 13 27; do FGB_GEMM_BAND_N=$bn FGB_GEMM_GROUP_M=$gm python tools/gemm_sweep.py ffn2,o; done; done
for bn in 56 28 14 7; do for gm in 4 8 13 27 54; do FGB_GEMM_BAND_N=$bn FGB_GEMM_GROUP_M=$gm python tools/gemm_sweep.py ffn1,qkv; done; done
