#!/usr/bin/env python
"""bench.py — the reference's headline metric on the reference's headline config (BASELINE.json):
Wan2.2-TI2V-5B denoise steps/s (and video tokens/s, % of bf16 tensor peak) at 704x1280x121, bf16, CFG on,
merged motion LoRA.  One "step" = one denoising step of WanVideoPipeline.__call__ (wan_video.py:285-309):
two DiT forwards (positive / negative prompt), CFG combine, flow-match Euler update, first-frame restore.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the video's token sequence is split across ranks
by Ulysses sequence parallelism (strong scaling: the job is one video).  `--impl reference` times the
reference's algorithm on the host CPU cores through the oracle port (a bounded sample, see cpu_sample()).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

HEIGHT, WIDTH, FRAMES = 704, 1280, 121
NUM_INFERENCE_STEPS, CFG_SCALE, SIGMA_SHIFT, TEXT_LEN, LORA_RANK = 50, 5.0, 5.0, 512, 32
METRIC, UNIT = "dit_denoise_steps_per_s", "steps/s"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (self-attention at S = 27 280, 24 heads)
    from the committed `ncu --set full` summary; None if the summary is missing.  Only meaningful for the 1-GPU shape."""
    path = os.path.join(REPO, "profiles", "r01_ncu_attn_summary.csv")
    try:
        tot = 0.0
        with open(path) as f:
            for line in f:
                parts = line.strip().split(",")
                if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(parts[1], None)
                    if scale is None:
                        return None
                    tot += float(parts[2]) * scale
        return tot or None
    except OSError:
        return None


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"], "hbm": p["hbm_gbs"],
                "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# CPU leg: the reference's algorithm (oracle port) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample(steps: int, warmup: int, sample_tokens: int = 4096):
    """Time ONE DiT block (fp32, real TI2V-5B dims, per-token modulation, 512 text tokens) at `sample_tokens`
    video tokens with all host threads, separately for self-attention (cost ~ S^2) and everything else
    (cost ~ S), and extrapolate both to S = 27 280, x30 blocks, x2 forwards per CFG step."""
    import torch

    from oracle import wan_dit_oracle as o

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = o.DiTConfig(num_layers=1)
    w = o.make_weights(cfg, seed=0)
    f, h, wd = 4, 16, sample_tokens // 64
    s = f * h * wd
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, s, cfg.dim, generator=g)
    ctx = torch.randn(1, TEXT_LEN, cfg.dim, generator=g)
    t_mod = torch.randn(1, s, 6, cfg.dim, generator=g) * 0.1
    freqs = o.rope_freqs(o.rope_tables_3d(cfg.head_dim), f, h, wd)
    attn_time = [0.0]

    def timed_attention(q, k, v, heads):
        t0 = time.perf_counter()
        out = o.attention(q, k, v, heads)
        attn_time[0] += time.perf_counter() - t0
        return out

    per_step = []
    with torch.no_grad():
        for i in range(warmup + steps):
            attn_time[0] = 0.0
            t0 = time.perf_counter()
            o.dit_block(w, 0, x, ctx, t_mod, freqs, cfg, attn_fn=timed_attention)
            total = time.perf_counter() - t0
            if i >= warmup:
                per_step.append((total, attn_time[0]))
    total = sorted(p[0] for p in per_step)[len(per_step) // 2]
    attn = sorted(p[1] for p in per_step)[len(per_step) // 2]
    s_full = (FRAMES - 1) // 4 + 1
    s_full = s_full * (HEIGHT // 32) * (WIDTH // 32)
    ratio = s_full / s
    block_full = (total - attn) * ratio + attn * ratio * ratio
    step_seconds = block_full * 30 * 2
    return {
        "value": 1.0 / step_seconds, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": (f"oracle port of DiTBlock.forward (fp32, torch CPU, {cores} threads): 1 of 30 blocks at {s} video tokens + "
                   f"{TEXT_LEN} text tokens, median of {steps} runs = {total:.3f} s (self-attention {attn:.3f} s); extrapolated "
                   f"to S={s_full} (attention x{ratio * ratio:.1f}, rest x{ratio:.2f}), x30 blocks, x2 CFG forwards"),
        "ms_per_sample": total * 1e3,
    }, step_seconds


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    base, step_seconds = cpu_sample(args.steps, args.warmup)
    s_full = ((FRAMES - 1) // 4 + 1) * (HEIGHT // 32) * (WIDTH // 32)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_seconds * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1), "video_tokens_per_s": 2 * s_full / step_seconds,
        "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus, layout=None):
    return {
        "workload": "Wan2.2-TI2V-5B denoise step, 704x1280x121 (latent 1x48x31x44x80, S=27280 video tokens, 512 text tokens), "
                    "bf16, CFG on (2 DiT forwards/step, cfg_scale 5), merged rank-32 motion LoRA, 50-step flow-match schedule (shift 5)",
        "tokens": 27280, "text_tokens": TEXT_LEN, "layers": 30, "parallelism": layout or ("single_gpu" if n_gpus == 1 else f"{n_gpus}_gpus"),
        "l2_policy": "working set per step (10 GB weights + 2 GB activations) >> 126 MB L2; no explicit flush needed",
    }


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic
    from fairygen_b200.profiling import ClockSampler, KernelTimer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sp = None
    par = None
    sp_ways = 1
    layout_name = "single_gpu"
    if world > 1:
        from fairygen_b200.cfg_parallel import Layout, ParallelContext

        dist.init_process_group("nccl", init_method="env://", device_id=dev)
        # one video: CFG pair (positive / negative prompt on disjoint halves of the box) x Ulysses inside each half,
        # or pure Ulysses over all ranks with --layout sp
        layout = Layout(world, 1, 1, world) if args.layout == "sp" else Layout.auto(world, 1, True, fg.TI2V_5B.num_heads)
        par = ParallelContext(layout, exchange=args.exchange)
        sp = par.sequence_parallel()
        sp_ways = layout.sp
        layout_name = f"cfg{layout.cfg}_x_ulysses_sp{layout.sp}_{args.exchange}"

    cfg = fg.TI2V_5B
    shape = synthetic.latent_shape(cfg, HEIGHT, WIDTH, FRAMES)
    tokens = shape[2] * (shape[3] // 2) * (shape[4] // 2)
    engine = fg.WanDiTEngine(cfg, dev, sp=sp)
    sd = synthetic.random_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16, lora_rank=LORA_RANK)
    engine.load_state_dict(sd)
    del sd
    lat_h, z0_h, cp_h, cn_h = synthetic.synthetic_inputs(cfg, shape, text_len=TEXT_LEN)
    den = fg.WanDenoiser(engine, NUM_INFERENCE_STEPS, CFG_SCALE, SIGMA_SHIFT,
                         cfg_group=par if (par is not None and par.layout.cfg > 1) else None)
    lat = lat_h.to(dev).contiguous()
    z0, cp, cn = z0_h.to(dev).contiguous(), cp_h.to(dev), cn_h.to(dev)
    lat[:, :, 0:1] = z0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- warm-up (also fills the context K/V cache, as the pipeline's first step does)
    for i in range(args.warmup):
        den.step(i % den.num_steps, lat, cp, cn, z0)
    barrier()

    # ---- timed region 1: device-resident inputs; per-kernel CUDA events on the launching stream
    timer = KernelTimer()
    engine.timer = timer
    engine.kernel_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            den.step((args.warmup + i) % den.num_steps, lat, cp, cn, z0)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = engine.kernel_launches + args.steps  # + one fused scheduler kernel per step
    kernels = timer.summary()
    engine.timer = None
    ms_per_step = ms_total / args.steps
    steps_per_s = 1e3 / ms_per_step

    # ---- timed region 2: end to end through the public API with HOST (pinned) buffers
    out_h = torch.empty_like(lat_h).pin_memory()
    ts_h = den.model_timesteps.pin_memory()
    h2d = lat_h.numel() * 2 + z0_h.numel() * 2 + cp_h.numel() * 2 + cn_h.numel() * 2
    d2h = out_h.numel() * 2
    barrier()
    e0.record()
    for i in range(args.steps):
        lat_d = lat_h.to(dev, non_blocking=True)
        z0_d = z0_h.to(dev, non_blocking=True)
        cp_d, cn_d = cp_h.to(dev, non_blocking=True), cn_h.to(dev, non_blocking=True)
        den.step((args.warmup + i) % den.num_steps, lat_d, cp_d, cn_d, z0_d)
        out_h.copy_(lat_d, non_blocking=True)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    finite = bool(torch.isfinite(out_h.float()).all())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = measured_peaks()
    flops_fwd = fg.counted_flops(cfg, tokens, TEXT_LEN)
    achieved_tflops = 2 * flops_fwd * steps_per_s / 1e12
    # dominant kernel: self-attention (53 % of the counted FLOPs). Algorithmic FLOPs per launch = 4 * S^2 * (heads*128) / world
    attn = kernels.get("attn_self")
    attn_flops = 4.0 * tokens * tokens * cfg.dim / sp_ways
    roofline = None
    if attn:
        a = attn_flops / (attn["avg_ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "attn_fwd_kernel (self-attention)", "achieved": a, "peak": peaks["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": a / peaks["bf16_sustained"], "traffic": ncu_traffic_bytes() if world == 1 else None,
                    "traffic_note": "DRAM bytes of one launch (ncu --set full, profiles/r01_ncu_attn_summary.csv); algorithmic bytes = q,k,v,o "
                                    "once = 4*S*D*2 = 670 MB",
                    "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "flops_per_launch": attn_flops, "avg_launch_ms": attn["avg_ms"], "launches": attn["launches"]}
    gemm_ms = sum(v["total_ms"] for k, v in kernels.items() if k.startswith("gemm"))
    fwd_per_rank = 2 * sp_ways / world   # DiT forwards each rank takes part in per step (1 under CFG-parallel)
    gemm_flops = 30 * (12 * tokens * cfg.dim ** 2 + 4 * tokens * cfg.dim * cfg.ffn_dim) / sp_ways * fwd_per_rank * args.steps
    breakdown = {k: round(v["total_ms"] / args.steps, 3) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["total_ms"])}
    # memory-bound kernels: algorithmic bytes per launch (one read + one write of the [rows, D] bf16 matrix = 2X/P)
    # over the live CUDA-event time of the launch, against the measured HBM copy bandwidth
    x_bytes = 2.0 * (-(-tokens // sp_ways)) * cfg.dim * 2
    hbm = {}
    for name in ("ln_modulate", "ln_affine", "rmsnorm_rope", "rmsnorm"):
        kt = kernels.get(name)
        if kt and kt["avg_ms"] > 0:
            gbs = x_bytes / (kt["avg_ms"] * 1e-3) / 1e9
            hbm[name] = {"achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3), "bytes_per_launch": x_bytes,
                         "avg_launch_us": round(kt["avg_ms"] * 1e3, 2)}

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_sample(3, 1)

    line = {
        "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(world, layout_name),
        "video_tokens_per_s": 2 * tokens * steps_per_s, "dit_forwards_per_s": 2 * steps_per_s,
        "achieved_tflops": achieved_tflops, "frac_of_bf16_peak_burst": achieved_tflops / (world * peaks["bf16_burst"]),
        "frac_of_bf16_peak_sustained": achieved_tflops / (world * peaks["bf16_sustained"]),
        "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "fairygen_b200.WanDenoiser.step on pinned host latents/contexts, result copied back to host"},
        "gpu_launches": launches, "roofline": roofline,
        "gemm": {"tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None, "ms_per_step": gemm_ms / args.steps},
        "kernel_ms_per_step": breakdown, "memory_bound_kernels": hbm, "clocks": clocks.summary(), "output_finite": finite,
    }
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--layout", choices=["auto", "sp"], default="auto",
                    help="N>1: auto = CFG pair x Ulysses SP (N/2 ways); sp = pure Ulysses SP over all N ranks")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p",
                    help="Ulysses exchange: p2p = NVLink peer stores from our own kernels (attention epilogue writes the owner's "
                         "buffer); nccl = pack / all-to-all / unpack")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
