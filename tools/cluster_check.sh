#!/bin/bash
cd "$(dirname "$0")/.."
export FGB_ATTN_CLUSTER=1
run() { echo "== $*"; timeout 120 tools/kcheck "$@" 2>&1 | grep -E "rel_l2|TFLOP|error|Error|lse" | head -4; }
run attn 256 128 1
run attn 128 512 2
run attn 300 200 2
run attn 48 32 2
run attn 1000 1000 3
run attn 1025 77 4
run attn 1000 3000 3
run attn 515 4100 5
run attn 27280 27280 24 8 0
FGB_ATTN_EMU=0 run attn 27280 27280 24 8 0
FGB_ATTN_EMU=2 run attn 27280 27280 24 8 0
run attn 27280 27280 6 8 0
run attn 27280 512 24 8 0
unset FGB_ATTN_CLUSTER
run attn 27280 27280 24 8 0
