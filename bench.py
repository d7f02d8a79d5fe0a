#!/usr/bin/env python
"""bench.py — the reference's headline metric on the reference's headline config (BASELINE.json):
Wan2.2-TI2V-5B denoise steps/s (and video tokens/s, % of bf16 tensor peak) at 704x1280x121, bf16, CFG on,
merged motion LoRA.  One "step" = one denoising step of WanVideoPipeline.__call__ (wan_video.py:285-309):
two DiT forwards (positive / negative prompt), CFG combine, flow-match Euler update, first-frame restore.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload headline|multishot|train|vae_encode]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the video's token sequence is split across ranks by Ulysses
sequence parallelism, optionally under a CFG pair (strong scaling: the job is one video).  Prints ONE JSON line on rank 0.

What the line carries beyond the contract keys:
  * N = 1: `gpu_reference` — the UNMODIFIED reference (baseline/_ref) in bf16 on the same B200 at the headline shape (forward
    time, rel-L2 of our forward against it) and the library kernels it lowers to, timed alone on the step's shapes: cuBLASLt per
    GEMM shape, flash-attn 2 / SDPA (cuDNN, flash) for self- and cross-attention, each beside our kernel;
    `attn_robustness` — how often the bounded-score fast path of the attention kernel applies (fallback CTA fraction), also with
    the q/k norm weights scaled up (--qk-norm-scale, default sweep 1 / 2.5 / 4).
  * N > 1: `parity` — 2 denoise steps of a RAGGED shape (480x832x81, S = 8190: 2 pad rows at SP 4 / 8) under the parallel layout
    against a single-rank engine and against the unmodified reference's bf16 loop on rank 0 (the run fails above 1e-2);
    `layouts` — steps/s of pure Ulysses SP over all N ranks beside CFG-pair x Ulysses (BASELINE config 3 as written).
  * `--impl reference`: the unmodified reference on the host cores at BASELINE config 1 (measured), scaled to the headline by
    counted FLOPs (labelled as extrapolated).
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

HEIGHT, WIDTH, FRAMES = 704, 1280, 121
if os.environ.get("FGB_BENCH_SHAPE"):   # experiment hook (e.g. 352x1280x121 at 2 GPUs = the per-rank GEMM shapes of SP4); the line says so
    HEIGHT, WIDTH, FRAMES = (int(v) for v in os.environ["FGB_BENCH_SHAPE"].split("x"))
PARITY_SHAPE = (480, 832, 81)   # ragged under SP 4 / 8: S = 8190
NUM_INFERENCE_STEPS, CFG_SCALE, SIGMA_SHIFT, TEXT_LEN, LORA_RANK = 50, 5.0, 5.0, 512, 32
METRIC, UNIT = "dit_denoise_steps_per_s", "steps/s"
PARITY_TOL = 1e-2


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (self-attention at S = 27 280, 24 heads)
    from the newest committed `ncu --set full` summary; None if there is none.  Only meaningful for the 1-GPU shape."""
    for name in ("r02final_ncu_attn_summary.csv", "r02_ncu_attn_summary.csv", "r01_ncu_attn_summary.csv"):
        path = os.path.join(REPO, "profiles", name)
        try:
            tot = 0.0
            with open(path) as f:
                for line in f:
                    parts = line.strip().split(",")
                    if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(parts[1], None)
                        if scale is None:
                            return None, None
                        tot += float(parts[2]) * scale
            if tot:
                return tot, "profiles/" + name
        except OSError:
            continue
    return None, None


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"], "hbm": p["hbm_gbs"],
                "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def headline_tokens():
    return ((FRAMES - 1) // 4 + 1) * (HEIGHT // 32) * (WIDTH // 32)


def workload_config(n_gpus, layout=None):
    return {
        "workload": "Wan2.2-TI2V-5B denoise step, 704x1280x121 (latent 1x48x31x44x80, S=27280 video tokens, 512 text tokens), "
                    "bf16, CFG on (2 DiT forwards/step, cfg_scale 5), merged rank-32 motion LoRA, 50-step flow-match schedule (shift 5)",
        "tokens": headline_tokens(), "shape_override": os.environ.get("FGB_BENCH_SHAPE"), "text_tokens": TEXT_LEN, "layers": 30, "parallelism": layout or ("single_gpu" if n_gpus == 1 else f"{n_gpus}_gpus"),
        "l2_policy": "working set per step (10 GB weights + 2 GB activations) >> 126 MB L2; no explicit flush needed",
    }


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on the host CPU cores (rank 0 only)
# ------------------------------------------------------------------------------------------------
def cpu_port_sample(steps: int, warmup: int, sample_tokens: int = 4096):
    """Fallback when baseline/_ref is not installed: ONE DiT block of the oracle port (fp32) at `sample_tokens` tokens,
    self-attention (~S^2) and the rest (~S) timed separately and extrapolated to S = 27 280, x30 blocks, x2 forwards."""
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
    ratio = headline_tokens() / s
    step_seconds = ((total - attn) * ratio + attn * ratio * ratio) * 30 * 2
    return {"value": 1.0 / step_seconds, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"oracle port of DiTBlock.forward (fp32, torch CPU, {cores} threads): 1 of 30 blocks at {s} video tokens + "
                       f"{TEXT_LEN} text tokens, median of {steps} runs = {total:.3f} s (self-attention {attn:.3f} s); extrapolated "
                       f"to S={headline_tokens()} (attention x{ratio * ratio:.1f}, rest x{ratio:.2f}), x30 blocks, x2 CFG forwards"),
            "extrapolated": {"seconds_per_step": step_seconds}}


def cpu_reference(steps: int, warmup: int, with_block_sample: bool = True):
    import fairygen_b200 as fg
    from baseline import ref_loader as rl

    if rl.available():
        from baseline import reference_arm as ra

        return ra.cpu_reference_run(fg, fg.TI2V_5B, steps, warmup, headline_tokens(), TEXT_LEN, with_block_sample)
    return cpu_port_sample(steps, warmup)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    base = cpu_reference(args.steps, args.warmup)
    step_seconds = base["extrapolated"]["seconds_per_step"]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_seconds * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "value_kind": "EXTRAPOLATED to the headline config from the measured sample (see cpu_baseline.measured / .extrapolated); "
                      "a headline-size fp32 CPU step takes hours",
        "config": workload_config(1), "video_tokens_per_s": 2 * headline_tokens() / step_seconds,
        "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    if "measured" in base:
        line["sample_ms_per_step"] = base["measured"]["seconds_per_forward"] * 1e3
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
class Harness:
    """One process of the job: device, process group, weights, headline inputs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import fairygen_b200 as fg
        from fairygen_b200 import synthetic

        self.torch, self.dist, self.fg, self.synthetic = torch, dist, fg, synthetic
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if args.gpus != self.world and self.world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"   # the version banner goes to stdout, in front of the one JSON line
            dist.init_process_group("nccl", init_method="env://", device_id=self.dev)
        self.cfg = fg.TI2V_5B
        self.sd = synthetic.random_state_dict(self.cfg, seed=0, device=self.dev, dtype=torch.bfloat16, lora_rank=LORA_RANK)
        self.base_engine = None   # the first engine packs the weights; later ones share the packed tensors

    def make_engine(self, sp=None):
        eng = self.fg.WanDiTEngine(self.cfg, self.dev, sp=sp)
        if self.base_engine is None:
            eng.load_state_dict(self.sd)
            self.base_engine = eng
        else:
            eng.share_weights_from(self.base_engine)
        return eng

    def make_parallel(self, layout_name):
        """(ParallelContext or None, SequenceParallel or None, label) for 'auto' (CFG pair x Ulysses) or 'sp' (pure Ulysses)."""
        if self.world == 1:
            return None, None, "single_gpu"
        from fairygen_b200.cfg_parallel import Layout, ParallelContext

        layout = Layout(self.world, 1, 1, self.world) if layout_name == "sp" else Layout.auto(self.world, 1, True, self.cfg.num_heads)
        par = ParallelContext(layout, exchange=self.args.exchange)
        return par, par.sequence_parallel(), f"cfg{layout.cfg}_x_ulysses_sp{layout.sp}_{self.args.exchange}"

    def make_denoiser(self, layout_name):
        par, sp, label = self.make_parallel(layout_name)
        eng = self.make_engine(sp)
        den = self.fg.WanDenoiser(eng, NUM_INFERENCE_STEPS, CFG_SCALE, SIGMA_SHIFT,
                                  cfg_group=par if (par is not None and par.layout.cfg > 1) else None)
        return den, eng, par, label

    def inputs(self, height, width, frames, pin=True):
        shape = self.synthetic.latent_shape(self.cfg, height, width, frames)
        return shape, self.synthetic.synthetic_inputs(self.cfg, shape, text_len=TEXT_LEN, pin=pin)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def time_steps(self, den, dev_inputs, first, count):
        """`count` denoise steps bracketed by barrier + synchronize, CUDA events, max over ranks -> ms per step."""
        torch = self.torch
        lat, z0, cp, cn = dev_inputs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for i in range(count):
            den.step((first + i) % den.num_steps, lat, cp, cn, z0)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / count


def attn_stats(engine):
    """(CTAs that ran the bounded-score fast path, CTAs that fell back to the running-max path) since the last reset."""
    from fairygen_b200 import ops

    return ops.attention_stats(engine.device)


def robustness_block(h: Harness, den, eng, dev_inputs, scales):
    """Step time and fallback fraction of the attention kernel with norm_q / norm_k weights multiplied by each scale
    (per-head score bound ~ 16.3 * scale^2 in log2 units): how the headline number depends on weight statistics."""
    from fairygen_b200 import ops

    out = {}
    for s in scales:
        for b in eng.blocks:
            b.nq.mul_(s)
            b.nk.mul_(s)
        try:
            _step(den, 0, _clone_lat(dev_inputs))          # warm
            ops.attention_stats_reset(eng.device)
            ms = h.time_steps(den, _clone_lat(dev_inputs), 1, 1)
            fast, fell = ops.attention_stats(eng.device)
            lat = _clone_lat(dev_inputs)[0]
            den.step(1, lat, dev_inputs[2], dev_inputs[3], dev_inputs[1])
            finite = bool(h.torch.isfinite(lat.float()).all())
        finally:
            for b in eng.blocks:
                b.nq.div_(s)
                b.nk.div_(s)
        out[f"x{s:g}"] = {"ms_per_step": round(ms, 2), "attn_ctas_fast_path": fast, "attn_ctas_fallback": fell,
                          "attn_fallback_frac": (fell / (fast + fell)) if (fast + fell) else None, "output_finite": finite}
    return out


def _clone_lat(dev_inputs):
    lat, z0, cp, cn = dev_inputs
    return lat.clone(), z0, cp, cn


def _step(den, index, dev_inputs):
    lat, z0, cp, cn = dev_inputs
    return den.step(index % den.num_steps, lat, cp, cn, z0)


def parity_block(h: Harness, den_par, eng_sp=None):
    """2 denoise steps at the ragged shape under the parallel layout vs (a) a plain single-rank engine on rank 0 and (b) the
    UNMODIFIED reference's bf16 denoise loop on rank 0's GPU (its own scheduler + model_fn).  Every rank takes part in the parallel
    run; rank 0 alone runs the comparisons while the others wait at the barrier."""
    torch = h.torch
    shape, (lat_h, z0_h, cp_h, cn_h) = h.inputs(*PARITY_SHAPE, pin=False)
    dev = h.dev
    lat, z0, cp, cn = lat_h.to(dev).contiguous(), z0_h.to(dev).contiguous(), cp_h.to(dev), cn_h.to(dev)
    lat[:, :, 0:1] = z0
    steps = range(2)
    lat_par = lat.clone()
    for i in steps:
        den_par.step(i, lat_par, cp, cn, z0)
    # one DiT forward under PURE Ulysses over all ranks (every rank takes part): the per-forward number of the north star
    ts0 = den_par.model_timesteps[0:1]
    fwd_sp = eng_sp.forward(lat, ts0, cp, True) if eng_sp is not None else None
    h.barrier()
    tokens = shape[2] * (shape[3] // 2) * (shape[4] // 2)
    out = {"shape": f"{PARITY_SHAPE[0]}x{PARITY_SHAPE[1]}x{PARITY_SHAPE[2]} (S={tokens})", "denoise_steps": len(steps), "tol": PARITY_TOL}
    if h.rank == 0:
        single = h.fg.WanDenoiser(h.make_engine(None), NUM_INFERENCE_STEPS, CFG_SCALE, SIGMA_SHIFT)
        lat_one = lat.clone()
        for i in steps:
            single.step(i, lat_one, cp, cn, z0)
        rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())  # noqa: E731
        out["rel_l2_vs_single"] = rel(lat_par, lat_one)
        from baseline import ref_loader as rl

        if rl.available():
            from baseline import reference_arm as ra

            lat_ref = ra.gpu_reference_denoise(h.cfg, h.sd, lat, z0, cp, cn, dev, steps, NUM_INFERENCE_STEPS, CFG_SCALE, SIGMA_SHIFT)
            out["rel_l2_vs_reference_bf16"] = rel(lat_par, lat_ref)
            out["rel_l2_single_vs_reference_bf16"] = rel(lat_one, lat_ref)
            if fwd_sp is not None:
                want, _, _ = ra.gpu_reference_forward(h.cfg, h.sd, lat, ts0.to(device=dev, dtype=torch.bfloat16), cp, dev, forwards=0)
                out["rel_l2_vs_reference_bf16_one_forward_pure_sp"] = rel(fwd_sp, want)
                out["pure_sp_ways"] = eng_sp.sp.world
            out["reference"] = "unmodified reference denoise loop body (baseline/_ref: model_fn_wan_video x2, CFG, FlowMatchScheduler.step) in bf16 on rank 0"
        else:
            out["rel_l2_vs_reference_bf16"] = None
            out["reference"] = "baseline/_ref not installed"
        worst = max(v for k, v in out.items() if k.startswith("rel_l2_vs") and isinstance(v, float))
        out["ok"] = bool(worst <= PARITY_TOL and torch.isfinite(lat_par.float()).all())
        del single
        torch.cuda.empty_cache()
    h.barrier()
    return out


def run_ours(args):
    h = Harness(args)
    torch, fg, dist = h.torch, h.fg, h.dist
    from fairygen_b200 import ops
    from fairygen_b200.profiling import ClockSampler, KernelTimer

    world, rank, dev, cfg = h.world, h.rank, h.dev, h.cfg
    den, engine, par, layout_name = h.make_denoiser(args.layout)
    sp_ways = par.layout.sp if par is not None else 1
    shape, (lat_h, z0_h, cp_h, cn_h) = h.inputs(HEIGHT, WIDTH, FRAMES)
    tokens = shape[2] * (shape[3] // 2) * (shape[4] // 2)
    lat = lat_h.to(dev).contiguous()
    z0, cp, cn = z0_h.to(dev).contiguous(), cp_h.to(dev), cn_h.to(dev)
    lat[:, :, 0:1] = z0
    dev_inputs = (lat, z0, cp, cn)
    if args.qk_norm_scale != 1.0:
        for b in engine.blocks:
            b.nq.mul_(args.qk_norm_scale)
            b.nk.mul_(args.qk_norm_scale)

    # ---- warm-up (also fills the context K/V cache, as the pipeline's first step does)
    for i in range(args.warmup):
        den.step(i % den.num_steps, lat, cp, cn, z0)
    h.barrier()

    # ---- timed region 1: device-resident inputs; per-kernel CUDA events on the launching stream
    timer = KernelTimer()
    engine.timer = timer
    engine.kernel_launches = 0
    ops.attention_stats_reset(dev)
    with ClockSampler(h.local_rank) as clocks:
        ms_per_step = h.time_steps(den, dev_inputs, args.warmup, args.steps)
    launches = engine.kernel_launches + args.steps  # + one fused scheduler kernel per step
    kernels = timer.summary()
    engine.timer = None
    rank_kernels = None
    if world > 1 and os.environ.get("FGB_BENCH_RANK_KERNELS"):     # diagnostic: every rank's per-kernel ms/step (rank 0's is the line's)
        mine = {k: round(v["total_ms"] / args.steps, 3) for k, v in kernels.items()}
        rank_kernels = [None] * world
        dist.all_gather_object(rank_kernels, mine)
    fast, fell = ops.attention_stats(dev)
    steps_per_s = 1e3 / ms_per_step

    # ---- timed region 2: end to end through the public API with HOST (pinned) buffers. Per step: the latents and the first-frame
    # latents go host -> device, the result comes back; the two prompt embeddings are handed over as host tensors on every step
    # as well, and the denoiser recognises them on the host (they are per-video constants, PIPE:404-417) instead of re-uploading.
    out_h = torch.empty_like(lat_h).pin_memory()
    h2d = lat_h.numel() * 2 + z0_h.numel() * 2
    d2h = out_h.numel() * 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    den.step(args.warmup % den.num_steps, lat_h.to(dev), cp_h, cn_h, z0_h.to(dev))   # uploads the contexts once (untimed, like the warm-up)
    h.barrier()
    e0.record()
    for i in range(args.steps):
        lat_d = lat_h.to(dev, non_blocking=True)
        z0_d = z0_h.to(dev, non_blocking=True)
        den.step((args.warmup + i) % den.num_steps, lat_d, cp_h, cn_h, z0_d)
        out_h.copy_(lat_d, non_blocking=True)
    e1.record()
    h.barrier()
    e2e_ms = h.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    finite = bool(torch.isfinite(out_h.float()).all())
    if getattr(engine, "sp", None) is not None:
        engine.sp.check()       # an exchange barrier that gave up on a peer invalidates the run: fail here, not in the numbers

    # ---- N > 1: the other layout, then parity at the ragged shape (every rank takes part)
    layouts = None
    parity = None
    if world > 1 and not args.no_extras:
        layouts = {layout_name: round(steps_per_s, 4)}
        other = "sp" if args.layout == "auto" else "auto"
        den2, eng2, par2, label2 = h.make_denoiser(other)
        if label2 != layout_name:
            inputs2 = _clone_lat(dev_inputs)
            for i in range(2):
                _step(den2, i, inputs2)
            ms2 = h.time_steps(den2, inputs2, 2, min(args.steps, 4))
            layouts[label2] = round(1e3 / ms2, 4)
        eng_sp = eng2 if args.layout == "auto" else engine      # the engine whose Ulysses group spans all ranks
        parity = parity_block(h, den, eng_sp if (eng_sp.sp is not None and eng_sp.sp.world == world) else None)
        del den2, eng2, eng_sp
        torch.cuda.empty_cache()
        flag = torch.tensor([1 if (rank != 0 or parity.get("ok", False)) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity_ok = bool(flag.item())
    else:
        parity_ok = True

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0 if parity_ok else 3

    peaks = measured_peaks()
    flops_fwd = fg.counted_flops(cfg, tokens, TEXT_LEN)
    achieved_tflops = 2 * flops_fwd * steps_per_s / 1e12
    # dominant kernel: self-attention (53 % of the counted FLOPs). Algorithmic FLOPs per launch = 4 * S^2 * (heads*128) / world
    attn = kernels.get("attn_self")
    attn_flops = 4.0 * tokens * tokens * cfg.dim / sp_ways
    roofline = None
    if attn:
        a = attn_flops / (attn["avg_ms"] * 1e-3) / 1e12
        traffic, traffic_src = ncu_traffic_bytes() if world == 1 else (None, None)
        roofline = {"bound": "tensor", "kernel": "attn_fwd_kernel (self-attention)", "achieved": a, "peak": peaks["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": a / peaks["bf16_sustained"], "traffic": traffic,
                    "traffic_note": f"DRAM bytes of one launch (ncu --set full, {traffic_src}); algorithmic bytes = q,k,v,o once = 4*S*D*2 = 670 MB",
                    "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "frac_of_burst_peak": a / peaks["bf16_burst"],
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
            # the q/k norm + RoPE launch (fgb_qk_norm_rope; fgb_recv_norm_rope under Ulysses) reads and writes BOTH q and k: 2 x 2X/P
            nbytes = 2 * x_bytes if name == "rmsnorm_rope" else x_bytes
            gbs = nbytes / (kt["avg_ms"] * 1e-3) / 1e9
            hbm[name] = {"achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3), "bytes_per_launch": nbytes,
                         "avg_launch_us": round(kt["avg_ms"] * 1e3, 2)}

    line = {
        "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(world, layout_name),
        "video_tokens_per_s": 2 * tokens * steps_per_s, "dit_forwards_per_s": 2 * steps_per_s,
        "achieved_tflops": achieved_tflops, "frac_of_bf16_peak_burst": achieved_tflops / (world * peaks["bf16_burst"]),
        "frac_of_bf16_peak_sustained": achieved_tflops / (world * peaks["bf16_sustained"]),
        "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "context_bytes_uploaded_once": cp_h.numel() * 2 + cn_h.numel() * 2,
                "api": "fairygen_b200.WanDenoiser.step: pinned host latents + first-frame latents copied in and the result copied back "
                       "every step; host prompt embeddings passed every step, recognised on the host (uploaded once per video)"},
        "gpu_launches": launches, "roofline": roofline,
        "gemm": {"tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None, "ms_per_step": gemm_ms / args.steps},
        "kernel_ms_per_step": breakdown, "memory_bound_kernels": hbm, "clocks": clocks.summary(), "output_finite": finite,
        "attn_fallback_frac": (fell / (fast + fell)) if (fast + fell) else None,
        "attn_ctas": {"fast_path": fast, "fallback": fell, "qk_norm_scale": args.qk_norm_scale},
    }
    if rank_kernels is not None:
        line["kernel_ms_per_step_by_rank"] = rank_kernels
    if layouts is not None:
        line["layouts"] = layouts
    if parity is not None:
        line["parity"] = parity
    if world == 1 and not args.no_extras:
        if args.qk_norm_scale == 1.0:
            line["attn_robustness"] = robustness_block(h, den, engine, dev_inputs, (2.5, 4.0))
        from baseline import reference_arm as ra

        ts = den.model_timesteps[args.warmup % den.num_steps: args.warmup % den.num_steps + 1]
        lat0 = lat_h.to(dev).contiguous()
        lat0[:, :, 0:1] = z0
        try:
            line["gpu_reference"] = ra.gpu_reference_block(fg, cfg, engine, h.sd, lat0, ts.to(device=dev, dtype=torch.bfloat16), cp, dev,
                                                           tokens, TEXT_LEN, steps_per_s)
        except Exception as e:   # the reference arm must never take the product's number down with it
            line["gpu_reference"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference(2, 1, with_block_sample=True)
        except Exception as e:
            line["cpu_baseline"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0 if parity_ok else 3


def run_other_workload(args):
    """BASELINE configs 4 / 5 and the VAE encoder through the same entry point (each tool prints its own JSON line)."""
    import runpy

    tool = {"multishot": "bench_multishot.py", "train": "bench_train.py", "vae_encode": "bench_vae.py"}[args.workload]
    argv = [tool, "--steps", str(args.steps), "--warmup", str(args.warmup)]
    if args.workload == "vae_encode":
        argv = [tool, "--encode", "--repeat", str(max(1, args.steps))]
    sys.argv = argv
    runpy.run_path(os.path.join(REPO, "tools", tool), run_name="__main__")
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["headline", "multishot", "train", "vae_encode"], default="headline",
                    help="headline = BASELINE configs 2-3 (the contract line); multishot = config 4 (tools/bench_multishot.py); "
                         "train = config 5 (tools/bench_train.py); vae_encode = the VAE38 encoder (tools/bench_vae.py --encode)")
    ap.add_argument("--layout", choices=["auto", "sp"], default="auto",
                    help="N>1: auto = CFG pair x Ulysses SP (N/2 ways); sp = pure Ulysses SP over all N ranks")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p",
                    help="Ulysses exchange: p2p = NVLink peer stores from our own kernels (attention epilogue writes the owner's "
                         "buffer); nccl = pack / all-to-all / unpack")
    ap.add_argument("--qk-norm-scale", type=float, default=1.0,
                    help="multiply every norm_q / norm_k weight (random init has ~1): larger per-head q/k norms push the "
                         "attention kernel's score bound up; the line reports the fallback fraction")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip gpu_reference / robustness (N=1) and layouts / parity (N>1)")
    args = ap.parse_args()
    if args.workload != "headline":
        return run_other_workload(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
