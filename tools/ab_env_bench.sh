#!/bin/bash
# A/B of runtime switches inside the real denoise step (power-capped clocks): tools/ab_env_bench.sh "VAR=a" "VAR=b" ...  (each twice, interleaved)
cd "$(dirname "$0")/.."
for pass in 1 2; do
  for setting in "$@"; do
    env $setting python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms_per_step']
print('$setting pass $pass |', round(d['ms_per_step'],1),'ms | clk', d['clocks']['sm_mhz'], '|', ' '.join(f'{n}={k[n]}' for n in ('attn_self','attn_cross','rmsnorm_rope','ln_modulate','ln_affine','rmsnorm') if n in k))"
  done
done
