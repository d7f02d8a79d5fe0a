cd /root/repo
run() { echo "== $*"; timeout 60 tools/kcheck "$@" 2>&1 | grep -E "rel_l2|TFLOP|error|Error" | head -3; }









run attn 27280 27280 24 5 0
FGB_ATTN_EMU=2 run attn 27280 27280 24 5 0
FGB_ATTN_EMU=3 run attn 27280 27280 24 5 0
run attn 27280 512 24 5 0
run attn 1000 1000 3
run attn 300 200 2
run attn 128 512 2
run attn 2048 2048 2
run attn 48 32 2
