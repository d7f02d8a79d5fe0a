#!/usr/bin/env python
"""One launch each of the library kernels the reference lowers to at the headline shapes (cuDNN SDPA self / cross attention,
cuBLASLt FFN1 / attn-o GEMMs) and of ours beside them — a short program to put under `ncu --set full`."""
import os
import sys

import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from fairygen_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
S, H, D = 27280, 24, 3072
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(S, D, device=dev, dtype=torch.bfloat16, generator=g)
k = torch.randn(S, D, device=dev, dtype=torch.bfloat16, generator=g)
v = torch.randn(S, D, device=dev, dtype=torch.bfloat16, generator=g)
o = torch.empty_like(q)
kmax2 = torch.zeros(H, device=dev, dtype=torch.float32)
qh, kh, vh = (t.view(-1, H, 128).transpose(0, 1).unsqueeze(0) for t in (q, k, v))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
        F.scaled_dot_product_attention(qh, kh, vh)
        F.scaled_dot_product_attention(qh, kh[:, :, :512], vh[:, :, :512])
    ops.head_norm_max(k, kmax2, H)
    ops.attention(q, k, v, o, H, kmax2=kmax2)
    ops.attention(q, k[:512], v[:512], o, H, kmax2=kmax2)
    a = q
    w1 = torch.randn(14336, D, device=dev, dtype=torch.bfloat16, generator=g) * 0.02
    b1 = torch.zeros(14336, device=dev, dtype=torch.bfloat16)
    c1 = torch.empty(S, 14336, device=dev, dtype=torch.bfloat16)
    F.linear(a, w1, b1)
    ops.gemm(a, w1, b1, c1)
    wo = w1[:D].contiguous()
    bo = b1[:D].contiguous()
    F.linear(a, wo, bo)
    ops.gemm(a, wo, bo, o)
torch.cuda.synchronize()
print("probe ok")
