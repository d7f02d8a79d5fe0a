#!/usr/bin/env python
"""Per-kernel times of single-GPU DiT forwards at an arbitrary video size (experiment harness, not the bench):
    python tools/step_probe.py --height 352 --width 640 --frames 121      # 6820 tokens: the GEMM shapes of a Ulysses SP4 rank
Prints one JSON line: ms per forward, per-kernel ms per forward and the TFLOP/s of each GEMM / attention kernel."""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=352)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--frames", type=int, default=121)
    ap.add_argument("--forwards", type=int, default=6)
    ap.add_argument("--gap-ms", type=float, default=0.0, help="host sleep between forwards (lets the power cap relax)")
    args = ap.parse_args()
    import time

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic
    from fairygen_b200.profiling import KernelTimer

    dev = torch.device("cuda", 0)
    cfg = fg.TI2V_5B
    sd = synthetic.random_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16, lora_rank=32)
    eng = fg.WanDiTEngine(cfg, dev)
    eng.load_state_dict(sd)
    shape = synthetic.latent_shape(cfg, args.height, args.width, args.frames)
    lat_h, z0_h, cp_h, cn_h = synthetic.synthetic_inputs(cfg, shape, text_len=512, pin=False)
    lat, cp = lat_h.to(dev).contiguous(), cp_h.to(dev)
    ts = torch.tensor([900.0])
    tokens = shape[2] * (shape[3] // 2) * (shape[4] // 2)
    for _ in range(3):
        eng.forward(lat, ts, cp, True)
    torch.cuda.synchronize()
    timer = KernelTimer()
    eng.timer = timer
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for _ in range(args.forwards):
        e0.record()
        eng.forward(lat, ts, cp, True)
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
        if args.gap_ms:
            time.sleep(args.gap_ms * 1e-3)
    k = {n: v["total_ms"] / args.forwards for n, v in timer.summary().items()}
    d, ffn = cfg.dim, cfg.ffn_dim
    fl = {"gemm_ffn1": 2 * tokens * d * ffn, "gemm_ffn2": 2 * tokens * d * ffn, "gemm_qkv": 6 * tokens * d * d, "gemm_o": 2 * tokens * d * d,
          "gemm_cross_q": 2 * tokens * d * d, "gemm_cross_o": 2 * tokens * d * d, "attn_self": 4 * tokens * tokens * d,
          "attn_cross": 4 * tokens * 512 * d}
    print(json.dumps({"tokens": tokens, "ms_per_forward": round(total / args.forwards, 3), "sk": os.environ.get("FGB_GEMM_SK", "1"),
                      "gap_ms": args.gap_ms,
                      "kernel_ms": {n: round(v, 3) for n, v in sorted(k.items(), key=lambda kv: -kv[1])},
                      "tflops": {n: round(cfg.num_layers * f / k[n] / 1e9) for n, f in fl.items() if n in k}}))


if __name__ == "__main__":
    main()
