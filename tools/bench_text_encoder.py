#!/usr/bin/env python
"""umT5-xxl text encoder at the size the pipeline runs it: 24 layers, dim 4096, 64 heads x 64, ffn 10240,
512-token prompts, positive + negative prompt of one call as ONE batch of 2 (random-init weights, synthetic ids).
Prints one JSON line: prompts/s with inputs resident (`value`), through the public call with host ids / mask and the
embedding read back (`e2e`), the share of the attention kernel, and the oracle on the host cores as `cpu_baseline`
(bounded sample: ONE layer of the 24, scaled)."""
import argparse
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--live", type=int, nargs=2, default=[96, 160], help="live tokens of the positive / negative prompt")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain launch sequence instead of the captured CUDA graph")
    args = ap.parse_args()
    from fairygen_b200 import ops, text_encoder as te

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = te.UMT5_XXL
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    for name, shape in te.param_shapes(cfg).items():
        if name.endswith("norm.weight") or "norm1" in name or "norm2" in name:
            t = 1 + 0.1 * torch.randn(shape, generator=g, device=dev)
        elif "pos_embedding" in name:
            t = 0.5 * torch.randn(shape, generator=g, device=dev)
        elif name == "token_embedding.weight":
            t = torch.randn(shape, generator=g, device=dev, dtype=torch.bfloat16)
        else:
            t = torch.randn(shape, generator=g, device=dev, dtype=torch.bfloat16) * (shape[1] ** -0.5) * (0.35 if ".attn.q." in name or ".attn.k." in name else 1.0)
        sd[name] = t.to(torch.bfloat16)
    enc = te.UMT5Encoder(cfg, dev, use_graph=not args.no_graph)
    enc.load_state_dict(sd)
    del sd
    B, L = 2, 512
    ids = torch.randint(2, cfg.vocab, (B, L), generator=torch.Generator().manual_seed(1))
    mask = torch.zeros(B, L, dtype=torch.long)
    for b, n in enumerate(args.live):
        mask[b, :n] = 1
        ids[b, n:] = 0
    ids_pin, mask_pin = ids.pin_memory(), mask.pin_memory()
    ids_dev, mask_dev = ids.to(dev), mask.to(dev)
    out_host = torch.empty(B, L, cfg.dim, dtype=torch.bfloat16).pin_memory()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for _ in range(args.warmup):
        enc.encode_prompts(ids_dev, mask_dev)
    enc.kernel_launches = 0
    ms = timed(lambda: enc.encode_prompts(ids_dev, mask_dev), args.steps)
    launches = enc.kernel_launches // args.steps

    def e2e():
        out_host.copy_(enc.encode_prompts(ids_pin, mask_pin), non_blocking=True)

    e2e()
    ms_e2e = timed(e2e, args.steps)
    # the attention kernel alone, on the buffers of the last layer (24 launches = one forward's worth)
    Le = max(args.live)
    ws = {"qkv": (torch.randn(B * Le, 3 * cfg.dim_attn, device=dev) * 0.3).to(torch.bfloat16),
          "o": torch.empty(B * Le, cfg.dim_attn, device=dev, dtype=torch.bfloat16)}
    bias = enc._bias_tables(Le)
    km = (mask_dev[:, :Le] != 0).to(torch.uint8).contiguous()
    da = cfg.dim_attn

    def attn_only():
        for i in range(cfg.num_layers):
            ops.t5_attention(ws["qkv"][:, :da], ws["qkv"][:, da:2 * da], ws["qkv"][:, 2 * da:], ws["o"], B, cfg.num_heads, bias=bias[i], key_mask=km)

    attn_only()
    ms_attn = timed(attn_only, args.steps)
    rows = B * max(args.live)     # encode_prompts computes the live prefix only
    gemm_flops = 2 * rows * cfg.num_layers * (4 * cfg.dim * cfg.dim_attn + 3 * cfg.dim * cfg.dim_ffn)
    weight_bytes = 2 * cfg.num_layers * (4 * cfg.dim * cfg.dim_attn + 3 * cfg.dim * cfg.dim_ffn)
    line = {
        "metric": "umt5_xxl_prompts_per_s", "value": B * 1e3 / ms, "unit": "prompts/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"umT5-xxl encoder, 24 layers, 2 prompts x 512 tokens in one batch (live {args.live}; padded tail not computed), random-init weights"},
        "e2e": {"value": B * 1e3 / ms_e2e, "unit": "prompts/s", "h2d_bytes_per_step": ids.numel() * 8 + mask.numel() * 8,
                "d2h_bytes_per_step": out_host.numel() * 2},
        "gpu_launches": launches, "cuda_graph": not args.no_graph, "attention_ms_per_step": ms_attn, "attention_share": ms_attn / ms,
        "gemm_tflops": gemm_flops / (ms * 1e-3) / 1e12, "weight_stream_gbps_lower_bound": weight_bytes / (ms * 1e-3) / 1e9,
    }
    if not args.no_cpu_baseline:
        from oracle import umt5_oracle as u
        ocfg = u.UMT5Config(vocab=1000, num_layers=1)
        w = u.make_weights(ocfg, seed=1)
        ids1 = ids % 1000
        torch.set_num_threads(os.cpu_count() or 1)
        u.encoder_forward(w, ocfg, ids1, mask)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            u.encoder_forward(w, ocfg, ids1, mask)
        per_layer = (time.perf_counter() - t0) / reps
        line["cpu_baseline"] = {"value": B / (per_layer * cfg.num_layers), "unit": "prompts/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"1 of {cfg.num_layers} layers (fp32 oracle, 2 x 512 tokens), time x {cfg.num_layers}"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
