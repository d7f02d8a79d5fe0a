#!/usr/bin/env python
"""Memory-bound kernels of the DiT block alone at the headline size (S = 27 280, D = 3072): GB/s of algorithmic bytes
against the measured HBM copy peak. Every launch touches > 300 MB, so nothing survives in the 126 MB L2 between launches."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from fairygen_b200 import ops  # noqa: E402

S, D = 27280, 3072
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(S, D, device=dev, generator=g).to(torch.bfloat16)
y = torch.empty_like(x)
qkv = torch.randn(S, 3 * D, device=dev, generator=g).to(torch.bfloat16)
vec = lambda: (1 + 0.1 * torch.randn(D, device=dev, generator=g)).to(torch.bfloat16)  # noqa: E731
w1, w2, w3, w4 = vec(), vec(), vec(), vec()
tab = torch.from_numpy(np.ascontiguousarray(ops.rope_table(128))).to(dev)
kmax = torch.zeros(24, dtype=torch.float32, device=dev)
grid = (31, 22, 40)
X = 2.0 * S * D * 2
cases = {
    "ln_modulate": (lambda: ops.ln_modulate(x, y, 1e-6, w1, w2, w3, w4, 880), X),
    "ln_affine": (lambda: ops.ln_affine(x, y, 1e-6, w1, w2), X),
    "rmsnorm": (lambda: ops.rmsnorm_rope(y, 1e-6, w1), X),
    "qk_norm_rope(+kmax)": (lambda: ops.qk_norm_rope(qkv, D, 1e-6, w1, w2, tab, grid, 0, kmax), 2 * X),
    "rmsnorm_rope(q) strided": (lambda: ops.rmsnorm_rope(qkv[:, :D], 1e-6, w1, tab, grid, 0), X),
    "head_norm_max": (lambda: ops.head_norm_max(qkv[:, D:2 * D], kmax, 24), X / 2),
    "torch copy": (lambda: y.copy_(x), X),
}
out = {}
for name, (fn, nbytes) in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    out[name] = {"us": round(us, 1), "GBps": round(nbytes / us / 1e3, 0)}
print(json.dumps(out))
