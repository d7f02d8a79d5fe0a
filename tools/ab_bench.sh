#!/bin/bash
# A/B the attention-kernel builds inside the real denoise step (sustained clocks): build_ab/lib_*.so vs the in-tree lib
cd "$(dirname "$0")/.."
for lib in "$@"; do
  echo "== $lib"
  if [ "$lib" = "intree" ]; then unset FGB_LIB_PATH; else export FGB_LIB_PATH=$PWD/build_ab/$lib; fi
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],4),'steps/s', round(d['ms_per_step'],1),'ms', 'attn_self',d['kernel_ms_per_step']['attn_self'], 'attn TF', round(d['roofline']['achieved'],1), 'clk', d['clocks']['sm_mhz'])"
done
