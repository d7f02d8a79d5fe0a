#!/bin/bash
# A/B of library builds on one box: tools/ab_libs.sh <what: attn|gemm> <seconds> lib1.so lib2.so ... ("intree" = the in-tree library)
# Each library is measured twice, interleaved, so box drift shows up as a difference between the two passes of one library.
cd "$(dirname "$0")/.."
what=$1; secs=$2; shift 2
for pass in 1 2; do
  for lib in "$@"; do
    if [ "$lib" = "intree" ]; then unset FGB_LIB_PATH; else export FGB_LIB_PATH=$PWD/build_ab/$lib; fi
    python tools/lib_compare.py --what $what --seconds $secs --ours-only 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    d=json.loads(line)
    vals=[(k,v['sustained_ms'],v['sustained_tflops']) for k,v in d.items() if k!='kernel']
    print('$lib pass $pass |', d['kernel'], '|', ' '.join(f'{ms:.4f}ms {tf:.0f}TF' for _,ms,tf in vals))"
  done
done
