#!/bin/bash
# One gpurun call that produces the round's profiling evidence (raw files under gpurun_out/, summarised into profiles/ by
# tools/ncu_summarise.py):  tools/profile_round.sh <tag>      e.g. r02
#   1. plain bench step (must exit 0) — the number that counts is never taken under ncu
#   2. ncu launch list of the same command (gpu__time_duration only; shares vs the bench line's live CUDA-event shares)
#   3. ncu --set full of ONE launch each: self-attention, cross-attention, FFN1 GEMM at the headline shapes (tools/kcheck)
cd "$(dirname "$0")/.."
tag=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $OUT/${tag}_prof_plain.json 2> $OUT/${tag}_prof_plain.err || { echo "plain bench failed"; tail -5 $OUT/${tag}_prof_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:fgb:: --csv \
  --log-file $OUT/${tag}_launches_raw.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $OUT/${tag}_prof_ncu_bench.log 2>&1
echo "launch list rc=$?"
full() {   # full <name> <kernel regex> <kcheck args...>
  local name=$1 regex=$2; shift 2
  KCHECK_BOUNDED=1 tools/kcheck "$@" 3 0 > $OUT/${tag}_kcheck_${name}.log 2>&1 || { echo "kcheck $name failed"; return; }
  KCHECK_BOUNDED=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$regex -s 2 -c 1 -f \
    -o $OUT/${tag}_ncu_${name} tools/kcheck "$@" 3 0 > $OUT/${tag}_ncu_${name}.log 2>&1
  echo "ncu $name rc=$?"
}
full attn attn_fwd attn 27280 27280 24
full attn_cross attn_fwd attn 27280 512 24
full gemm_ffn1 gemm_pair gemm 27280 14336 3072 1
full gemm_ffn2 gemm_pair gemm 27280 3072 14336 2
ls -la $OUT | grep ${tag}_ncu
