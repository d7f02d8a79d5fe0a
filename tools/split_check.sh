#!/bin/bash
# key-split tail of fgb_attn_fwd_ex: correctness on small ragged shapes, then timing with / without the split
cd "$(dirname "$0")/.."
run() { echo "== $*"; timeout 120 tools/kcheck "$@" 2>&1 | grep -E "rel_l2|TFLOP|error|Error|lse|workspace" | head -5; }
run attn 1000 3000 3
run attn 300 2100 2
run attn 515 4100 5
run attn 256 2048 1
run attn 27280 27280 3 5 1
KCHECK_NOSPLIT=1 run attn 27280 27280 3 5 0
run attn 27280 27280 6 5 0
KCHECK_NOSPLIT=1 run attn 27280 27280 6 5 0
run attn 27280 27280 12 5 0
KCHECK_NOSPLIT=1 run attn 27280 27280 12 5 0
run attn 27280 27280 24 5 0
KCHECK_NOSPLIT=1 run attn 27280 27280 24 5 0
run attn 27280 512 24 5 0
