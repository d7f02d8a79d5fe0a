#!/usr/bin/env python
"""Sustained TFLOP/s of fgb_gemm_bf16 on the step's GEMM shapes for the supertile given in the environment
(FGB_GEMM_GROUP_M / FGB_GEMM_BAND_N, read once per process).  Tuning harness: tools/gemm_sweep.sh loops over it."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from fairygen_b200 import ops  # noqa: E402

S = int(os.environ.get('ROWS', '27280'))
shapes = {"qkv": (9216, 3072), "o": (3072, 3072), "ffn1": (14336, 3072), "ffn2": (3072, 14336)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else list(shapes)
epi = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
out = []
for name in which:
    n, k = shapes[name]
    a = torch.randn(S, k, device=dev, dtype=torch.bfloat16, generator=g)
    w = torch.randn(n, k, device=dev, dtype=torch.bfloat16, generator=g) * k ** -0.5
    b = torch.randn(n, device=dev, dtype=torch.bfloat16, generator=g)
    c = torch.zeros(S, n, device=dev, dtype=torch.bfloat16)
    g0 = torch.randn(n, device=dev, dtype=torch.bfloat16, generator=g)
    sk = ops.gemm_workspace(dev) if os.environ.get("SK", "0") != "0" else None     # SK=1: stream-K tail (fgb_gemm_bf16_sk)
    fn = lambda: ops.gemm(a, w, b, c, epi, g0, g0, 880, sk_ws=sk)  # noqa: E731
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    iters = max(20, int(0.5 / (2.0 * S * n * k / 1.3e15)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    out.append(f"{name} {2.0 * S * n * k / ms / 1e9:.0f}")
print(f"rows={S} SK={os.environ.get('SK', '0')} GM={os.environ.get('FGB_GEMM_GROUP_M', '-')} BAND={os.environ.get('FGB_GEMM_BAND_N', '-')} BN={os.environ.get('FGB_GEMM_BN', '-')} lib={os.path.basename(os.environ.get('FGB_LIB_PATH', 'intree'))} epi={epi}: " + "  ".join(out))
