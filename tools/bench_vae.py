#!/usr/bin/env python
"""VAE38 decode at the headline size: 31 latent frames of 44 x 80 (704x1280x121 video), tiled with the
pipeline's defaults (tile 30 x 52, stride 15 x 26: 6 windows), random-init weights of the full-width decoder.  One JSON line:
seconds per video with the latents resident, through the public call with host latents and the video read back (`e2e`),
convolution FLOPs actually issued (grid borders and channel padding included) and the rate they ran at."""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _random_sd(shapes, dev):
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    for name, shape in shapes.items():
        if name.endswith("gamma"):
            t = 1 + 0.1 * torch.randn(shape, generator=g, device=dev)
        elif name.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g, device=dev)
        else:
            fan = 1
            for v in shape[1:]:
                fan *= v
            t = torch.randn(shape, generator=g, device=dev) * fan ** -0.5
        sd[name] = t
    return sd


def bench_encode(args):
    """VAE38 encoder (wan_video_vae.py:1103-1253 via WanVideoVAE.encode, called at PIPE:495 for the first-frame image and by the
    training data path for whole clips): full-width random-init encoder on (a) the 704x1280 first-frame image of the headline
    video, tiled as the pipeline does, and (b) the 49-frame 480x832 training clip of BASELINE config 5.  One JSON line."""
    from fairygen_b200 import vae, vae_encode

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    enc = vae_encode.VAE38Encoder(vae.VAE38, dev)
    enc.load_state_dict(_random_sd(vae_encode.enc_param_shapes(vae.VAE38), dev))
    cases = {"first_frame_image_704x1280": (3, 1, 704, 1280), "training_clip_49x480x832": (3, 49, 480, 832)}
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, shape in cases.items():
        host = torch.tanh(torch.randn(shape, generator=torch.Generator().manual_seed(2))).to(torch.bfloat16).pin_memory()
        devv = host.to(dev)
        run = lambda v: enc.encode([v], tiled=True, tile_size=(34, 34), tile_stride=(18, 16))   # noqa: E731  the reference's defaults
        z = run(devv)
        torch.cuda.synchronize()
        enc.kernel_launches = 0
        e0.record()
        for _ in range(args.repeat):
            z = run(devv)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.repeat
        launches = enc.kernel_launches // args.repeat
        z_host = torch.empty(z.shape, dtype=z.dtype).pin_memory()
        e0.record()
        for _ in range(args.repeat):
            z_host.copy_(run(host.to(dev, non_blocking=True)), non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / args.repeat
        out[name] = {"ms": round(ms, 2), "ms_e2e_host_in_host_out": round(ms_e2e, 2), "latent_shape": list(z.shape), "gpu_launches": launches,
                     "input_mpix_per_s": round(shape[1] * shape[2] * shape[3] / ms / 1e3, 1), "finite": bool(torch.isfinite(z.float()).all())}
    img = out["first_frame_image_704x1280"]
    print(json.dumps({
        "metric": "vae38_encode_images_per_s", "value": 1e3 / img["ms"], "unit": "images/s", "n_gpus": 1, "steps": args.repeat, "warmup": 1,
        "ms_per_step": img["ms"], "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Wan2.2 VAE38 tiled encode (tile 34x34, stride 18x16), full-width encoder, random-init weights"},
        "e2e": {"value": 1e3 / img["ms_e2e_host_in_host_out"], "unit": "images/s", "h2d_bytes_per_step": 3 * 704 * 1280 * 2,
                "d2h_bytes_per_step": 48 * 44 * 80 * 2},
        "gpu_launches": img["gpu_launches"], "cases": out, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=31, help="latent frames (31 = 121 video frames)")
    ap.add_argument("--height", type=int, default=44)
    ap.add_argument("--width", type=int, default=80)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--encode", action="store_true", help="time the VAE38 ENCODER instead (PIPE:490-497 first-frame image; training clip)")
    ap.add_argument("--repeat", type=int, default=3, help="--encode: timed repetitions per input")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel CUDA time of one decode (torch.profiler / CUPTI) to stderr")
    ap.add_argument("--torch-chain", action="store_true",
                    help="also time the reference's own bf16 op chain (torch / cuDNN kernels, the oracle functions run in bf16) on the "
                         "same windows on this GPU — the baseline leg; blending and the reference's per-tile CPU round trip excluded")
    args = ap.parse_args()
    if args.encode:
        return bench_encode(args)
    from fairygen_b200 import ops, vae

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = vae.VAE38
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    for name, shape in vae.param_shapes(cfg).items():
        if name.endswith("gamma"):
            t = 1 + 0.1 * torch.randn(shape, generator=g, device=dev)
        elif name.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g, device=dev)
        else:
            fan = 1
            for v in shape[1:]:
                fan *= v
            t = torch.randn(shape, generator=g, device=dev) * fan ** -0.5
        sd[name] = t
    dec = vae.VAE38Decoder(cfg, dev)
    dec.load_state_dict(sd)
    chain_w = {k: v.to(torch.bfloat16) for k, v in sd.items()} if args.torch_chain else None
    del sd
    z_host = torch.randn(1, 48, args.frames, args.height, args.width, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).pin_memory()
    z_dev = z_host.to(dev)
    tile, stride = (30, 52), (15, 26)
    flops = [0]
    orig = ops.conv_taps

    def counted(x, a_row0, w, bias, out, taps, grid_hw=(0, 0), epilogue=0):
        flops[0] += 2 * out.shape[0] * w.shape[0] * w.shape[1]
        return orig(x, a_row0, w, bias, out, taps, grid_hw, epilogue)

    ops.conv_taps = counted

    def run(z):
        return dec.decode(z, tiled=True, tile_size=tile, tile_stride=stride)

    for _ in range(args.warmup):
        run(z_dev)
    torch.cuda.synchronize()
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run(z_dev)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60), file=sys.stderr)
    flops[0] = 0
    dec.kernel_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = run(z_dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    conv_flops = flops[0] / args.steps
    launches = dec.kernel_launches // args.steps
    host_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    e0.record()
    for _ in range(args.steps):
        host_out.copy_(run(z_host), non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    chain = None
    if args.torch_chain:
        from oracle import vae38_oracle as o    # baseline leg only: the reference's op chain, timed, never part of the product path
        wins = vae.tile_tasks(args.height, args.width, tile, stride)
        with torch.no_grad():
            o.model_decode(chain_w, o.VAE38, z_dev[:, :, :2, :8, :8])       # warm cuDNN up
            torch.cuda.synchronize()
            e0.record()
            for h0, h1, w0, w1 in wins:
                o.model_decode(chain_w, o.VAE38, z_dev[:, :, :, h0:h1, w0:w1])
            e1.record()
            torch.cuda.synchronize()
        chain = {"ms_per_video": e0.elapsed_time(e1), "what": "reference op chain in bf16 via torch / cuDNN on the same 6 windows, same GPU; "
                 "no blending, no per-tile CPU round trip"}
    print(json.dumps({
        "metric": "vae38_decode_videos_per_s", "value": 1e3 / ms, "unit": "videos/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"Wan2.2 VAE38 tiled decode, latents 48x{args.frames}x{args.height}x{args.width} -> "
                               f"{4 * args.frames - 3} frames of {16 * args.height}x{16 * args.width}, tile {tile} stride {stride}, "
                               "full-width decoder (1024/512/256 channels), random-init weights"},
        "e2e": {"value": 1e3 / ms_e2e, "unit": "videos/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": z_host.numel() * 2,
                "d2h_bytes_per_step": host_out.numel() * host_out.element_size()},
        "gpu_launches": launches, "conv_flops_issued": conv_flops, "conv_tflops": conv_flops / (ms * 1e-3) / 1e12,
        "torch_bf16_chain": chain,
        "output_finite": bool(torch.isfinite(out.float()).all()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
    }))


if __name__ == "__main__":
    main()
