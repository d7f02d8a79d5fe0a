#!/bin/bash
# run a kcheck command for a while and sample SM clock / power meanwhile: tools/clock_probe.sh <env assignments...> -- <kcheck args>
cd "$(dirname "$0")/.."
envs=()
while [ "$1" != "--" ]; do envs+=("$1"); shift; done
shift
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap --format=csv,noheader -lms 100 > /tmp/clk.log &
SMI=$!
env "${envs[@]}" tools/kcheck "$@" 2>&1 | grep -E "TFLOP|error"
kill $SMI
sort /tmp/clk.log | uniq -c | sort -rn | head -4
