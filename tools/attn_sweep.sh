#!/bin/bash
# sweep the MUFU/FMA split of the attention softmax (FGB_ATTN_EMU of every 8 pairs emulated)
cd "$(dirname "$0")/.."
for e in ${EMUS:-0 1 2 3 4}; do
  echo "== EMU=$e"; FGB_ATTN_EMU=$e timeout 60 tools/kcheck attn 1000 1000 3 2>&1 | grep rel_l2
  FGB_ATTN_EMU=$e timeout 60 tools/kcheck attn 27280 27280 24 5 0 2>&1 | grep TFLOP
done
for a in "256 128 1" "300 200 2" "128 512 2" "2048 2048 2"; do timeout 60 tools/kcheck attn $a | grep rel_l2; done
timeout 60 tools/kcheck attn 27280 512 24 5 0 | grep TFLOP
