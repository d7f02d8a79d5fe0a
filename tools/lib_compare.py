#!/usr/bin/env python
"""Library kernels beside ours at the headline step's shapes, isolated AND sustained (back-to-back for ~1 s each, so the
1 kW power cap settles the clock the way it does inside a denoise step).  Experiment harness (not the bench):
    python tools/lib_compare.py [--what attn,gemm] [--seconds 1.0]
Prints one JSON object per kernel pair."""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def sustained(fn, seconds):
    """ms per call over a back-to-back run of ~`seconds`, plus the first-10 average (cold clocks)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    first = e0.elapsed_time(e1) / 10
    n = max(10, int(seconds * 1e3 / first))
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return {"first10_ms": round(first, 4), "sustained_ms": round(e0.elapsed_time(e1) / n, 4), "calls": n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="attn,gemm")
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--tokens", type=int, default=27280)
    ap.add_argument("--ours-only", action="store_true", help="A/B of library builds (FGB_LIB_PATH): skip the vendor kernels")
    args = ap.parse_args()
    from torch.nn.attention import SDPBackend, sdpa_kernel

    from fairygen_b200 import ops

    dev = torch.device("cuda", 0)
    S, H, D = args.tokens, 24, 3072
    g = torch.Generator(device=dev).manual_seed(0)
    if "attn" in args.what:
        for s_kv in (S, 512):
            q = torch.randn(S, D, device=dev, dtype=torch.bfloat16, generator=g)
            k = torch.randn(s_kv, D, device=dev, dtype=torch.bfloat16, generator=g)
            v = torch.randn(s_kv, D, device=dev, dtype=torch.bfloat16, generator=g)
            o = torch.empty_like(q)
            kmax2 = torch.zeros(H, device=dev, dtype=torch.float32)
            ops.head_norm_max(k, kmax2, H)
            qmax2 = torch.zeros(H, device=dev, dtype=torch.float32)
            ops.head_norm_max(q, qmax2, H)
            use_q = os.environ.get("QMAX", "1") != "0"      # QMAX=0: per-row bounds only (A/B of the head-level bound)
            qh, kh, vh = (t.view(-1, H, 128).transpose(0, 1).unsqueeze(0) for t in (q, k, v))
            fl = 4.0 * S * s_kv * D

            def ours():
                ops.attention(q, k, v, o, H, kmax2=kmax2, qmax2=qmax2 if use_q else None)

            def ours_runmax():
                ops.attention(q, k, v, o, H)

            def cudnn():
                with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
                    return F.scaled_dot_product_attention(qh, kh, vh)

            res = {"kernel": f"attention S_q={S} S_kv={s_kv}"}
            for rnd in range(2):   # twice, alternating, so neither side always runs on the hotter chip
                impls = (("ours_bounded", ours),) if args.ours_only else (("ours_bounded", ours), ("cudnn_sdpa", cudnn), ("ours_running_max", ours_runmax))
                for name, fn in impls:
                    r = sustained(fn, args.seconds)
                    r["sustained_tflops"] = round(fl / r["sustained_ms"] / 1e9, 1)
                    r["first10_tflops"] = round(fl / r["first10_ms"] / 1e9, 1)
                    res[f"{name}_{rnd}"] = r
                    time.sleep(0.5)
            print(json.dumps(res))
    if "gemm" in args.what:
        for name, n, kk in (("qkv", 9216, 3072), ("o", 3072, 3072), ("ffn1", 14336, 3072), ("ffn2", 3072, 14336)):
            a = torch.randn(S, kk, device=dev, dtype=torch.bfloat16, generator=g)
            w = torch.randn(n, kk, device=dev, dtype=torch.bfloat16, generator=g) * kk ** -0.5
            b = torch.randn(n, device=dev, dtype=torch.bfloat16, generator=g)
            c = torch.empty(S, n, device=dev, dtype=torch.bfloat16)
            fl = 2.0 * S * n * kk
            res = {"kernel": f"gemm {name} {S}x{n}x{kk}"}
            for rnd in range(2):
                impls = (("ours", lambda: ops.gemm(a, w, b, c)),) if args.ours_only else (("ours", lambda: ops.gemm(a, w, b, c)), ("cublas", lambda: F.linear(a, w, b)))
                for lab, fn in impls:
                    r = sustained(fn, args.seconds)
                    r["sustained_tflops"] = round(fl / r["sustained_ms"] / 1e9, 1)
                    r["first10_tflops"] = round(fl / r["first10_ms"] / 1e9, 1)
                    res[f"{lab}_{rnd}"] = r
                    time.sleep(0.5)
            print(json.dumps(res))


if __name__ == "__main__":
    main()
