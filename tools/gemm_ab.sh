#!/bin/bash
# A/B GEMM builds (build_ab/lib_*.so) on the four headline GEMM shapes with the stand-alone harness (LD_LIBRARY_PATH swap)
cd "$(dirname "$0")/.."
for lib in "$@"; do
  echo "== $lib"
  if [ "$lib" != "intree" ]; then mkdir -p /tmp/ab_$lib && cp build_ab/$lib /tmp/ab_$lib/libfairygen_b200.so; export LD_LIBRARY_PATH=/tmp/ab_$lib; else unset LD_LIBRARY_PATH; fi
  for s in "27280 3072 3072 2" "27280 9216 3072 0" "27280 14336 3072 1" "27280 3072 14336 2" "6820 3072 3072 2" "6820 14336 3072 1"; do
    tools/kcheck gemm $s 30 2>&1 | grep TFLOP
  done
done
