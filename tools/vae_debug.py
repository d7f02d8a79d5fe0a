#!/usr/bin/env python
"""Layer-by-layer comparison of the VAE38 decoder kernels with the oracle (reduced widths): prints the relative L2 error of every
intermediate grid (interior), and the largest value found on grid borders / padding channels (must be 0)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    from fairygen_b200 import ops, vae
    from oracle import vae38_oracle as o

    torch.cuda.set_device(0)
    ocfg = o.TINY
    cfg = vae.VAE38Config(z_dim=ocfg.z_dim, dec_dim=ocfg.dec_dim)
    w = o.make_weights(ocfg, seed=0)
    dec = vae.VAE38Decoder(cfg, "cuda")
    dec.load_state_dict(w)
    z = torch.randn((1, ocfg.z_dim, 3, 3, 4), generator=torch.Generator().manual_seed(1))
    got = []

    def trace(name, rows, T, h, w_, C):
        g = rows.float().view(T, h + 2, w_ + 2, -1)
        inner = g[:, 1:-1, 1:-1, :C].permute(3, 0, 1, 2).contiguous()
        edge = g.clone()
        edge[:, 1:-1, 1:-1, :C] = 0
        got.append((name, inner, float(edge.abs().max())))

    dec.trace = trace
    out = dec.decode(z.to(torch.bfloat16), tiled=False)
    ops.sync_check()
    want = []
    w16 = {k: v.to(torch.bfloat16).float().cuda() for k, v in w.items()}
    with torch.no_grad():
        ref = o.model_decode(w16, ocfg, z.to(torch.bfloat16).float().cuda(), trace=lambda n, t: want.append((n, t[0])))
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))  # noqa: E731
    print(f"{len(got)} traced grids (oracle: {len(want)})")
    for (n1, a, edge), (n2, b) in zip(got, want):
        flag = "" if (n1 == n2 and tuple(a.shape) == tuple(b.shape)) else f"  <-- MISMATCH {n2} {tuple(b.shape)}"
        err = rel(a, b) if tuple(a.shape) == tuple(b.shape) else float("nan")
        print(f"{n1:14s} {str(tuple(a.shape)):22s} rel_l2 {err:.3e}  border/pad max {edge:.2e}{flag}")
    print("final (clamped) rel_l2", rel(out[0].float(), ref[0].clamp(-1, 1)), tuple(out.shape), "finite", bool(torch.isfinite(out.float()).all()))


if __name__ == "__main__":
    main()
