#!/bin/bash
# A/B of runtime switches of ONE library build: tools/ab_env.sh <what> <seconds> "VAR=a" "VAR=b" ...  (each twice, interleaved)
cd "$(dirname "$0")/.."
what=$1; secs=$2; shift 2
for pass in 1 2; do
  for setting in "$@"; do
    env $setting python tools/lib_compare.py --what $what --seconds $secs --ours-only 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    d=json.loads(line)
    vals=[(k,v['sustained_ms'],v['sustained_tflops']) for k,v in d.items() if k!='kernel']
    print('$setting pass $pass |', d['kernel'], '|', ' '.join(f'{ms:.4f}ms {tf:.0f}TF' for _,ms,tf in vals))"
  done
done
