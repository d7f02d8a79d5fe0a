"""The reference arms of ``bench.py``: the UNMODIFIED reference (``baseline/_ref``, see install_ref.py / ref_loader.py)

  * on the host CPU cores (``--impl reference`` and the ``cpu_baseline`` block of our own line): BASELINE.json config 1 as
    written — Wan2.2-TI2V-5B random-init, ONE denoising forward, 17-frame 256x256 clip (latent 1x48x5x16x16, S = 320 video
    tokens), 512 text tokens, no CFG, fp32 — through ``model_fn_wan_video`` (PIPE:1122-1388), all host threads;
  * on the GPU (``gpu_reference`` block of the N = 1 line): the same reference code in bf16 on the B200 at the headline shape,
    plus the library kernels it lowers to, timed alone on the step's shapes — cuBLASLt (``F.linear``, DIT:140-146, 176-185,
    208-209) and flash-attn 2 / SDPA (``flash_attention``, DIT:27-60).

Nothing of ours is on the reference paths: the models are the reference's ``WanModel``, the calls are its own functions.
Measurement infrastructure only (never imported by ``fairygen_b200``).
"""
from __future__ import annotations

import os
import time
from statistics import median
from typing import Dict, Optional

CONFIG1 = dict(height=256, width=256, frames=17, text_len=512)   # BASELINE.json configs[0]


def _sync(dev):
    import torch

    if dev.type == "cuda":
        torch.cuda.synchronize(dev)


# ------------------------------------------------------------------------------------------------
# CPU: BASELINE config 1 with the unmodified reference
# ------------------------------------------------------------------------------------------------
class CpuReference:
    """Builds the reference WanModel (TI2V-5B dims, fp32, CPU) once; `forward_seconds()` times one config-1 forward."""

    def __init__(self, cfg):
        import torch

        from . import ref_loader as rl

        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.ref = rl.load()
        rl.select_attention_backend("cpu")
        t0 = time.perf_counter()
        self.cfg = cfg
        self.dit = rl.build_wan_model(cfg, device="cpu", dtype=torch.float32, fill="tiled")
        self.build_s = time.perf_counter() - t0
        g = torch.Generator().manual_seed(1)
        f = (CONFIG1["frames"] - 1) // 4 + 1
        self.latents = torch.randn(1, cfg.in_dim, f, CONFIG1["height"] // 16, CONFIG1["width"] // 16, generator=g)
        self.context = torch.randn(1, CONFIG1["text_len"], cfg.text_dim, generator=g)
        self.context[:, 64:] = 0
        self.timestep = torch.tensor([1000.0])
        self.tokens = f * (CONFIG1["height"] // 32) * (CONFIG1["width"] // 32)

    def forward_seconds(self) -> float:
        import torch

        with torch.no_grad():
            t0 = time.perf_counter()
            out = self.ref.wv.model_fn_wan_video(dit=self.dit, latents=self.latents, timestep=self.timestep, context=self.context,
                                                 fuse_vae_embedding_in_latents=True)
            dt = time.perf_counter() - t0
        assert out.shape == self.latents.shape and bool(torch.isfinite(out).all())
        return dt

    def block_sample(self, tokens: int = 4096, runs: int = 3):
        """One reference DiTBlock (DIT:195-229) at `tokens` video tokens: (seconds per block, seconds inside the self-attention's
        AttentionModule), measured with forward hooks — the S^2 share needed to extrapolate to 27 280 tokens."""
        import torch

        cfg, wd = self.cfg, self.ref.wd
        block = self.dit.blocks[0]
        f, h = 4, 16
        w = tokens // (f * h)
        s = f * h * w
        g = torch.Generator().manual_seed(2)
        x = torch.randn(1, s, cfg.dim, generator=g)
        ctx = torch.randn(1, CONFIG1["text_len"], cfg.dim, generator=g)
        t_mod = torch.randn(1, s, 6, cfg.dim, generator=g) * 0.1
        fr = self.dit.freqs
        freqs = torch.cat([fr[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), fr[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                           fr[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(s, 1, -1)
        acc = [0.0, 0.0]

        def pre(mod, args):
            acc[1] = time.perf_counter()

        def post(mod, args, out):
            acc[0] += time.perf_counter() - acc[1]

        h0 = block.self_attn.attn.register_forward_pre_hook(pre)
        h1 = block.self_attn.attn.register_forward_hook(post)
        totals, attns = [], []
        try:
            with torch.no_grad():
                for i in range(runs + 1):
                    acc[0] = 0.0
                    t0 = time.perf_counter()
                    block(x, ctx, t_mod, freqs)
                    dt = time.perf_counter() - t0
                    if i > 0:
                        totals.append(dt)
                        attns.append(acc[0])
        finally:
            h0.remove()
            h1.remove()
        del wd
        return s, median(totals), median(attns)


def extrapolate_step_seconds(fg, cfg, fwd_s: float, tokens_sample: int, tokens_full: int, text_len: int, forwards_per_step: int = 2):
    """Headline CFG-step time from a measured config-1 forward, scaled by the counted FLOPs of SURVEY.md §8(d) — the survey's
    rule. (At S = 320 the forward is bound by streaming 20 GB of fp32 weights, so this can over-state the CPU's time at
    S = 27 280; cpu_reference_run therefore also extrapolates from a compute-bound block sample and keeps the faster one.)"""
    ratio = fg.counted_flops(cfg, tokens_full, text_len) / fg.counted_flops(cfg, tokens_sample, text_len)
    return fwd_s * ratio * forwards_per_step, ratio


def cpu_reference_run(fg, cfg, steps: int, warmup: int, tokens_full: int, text_len: int, with_block_sample: bool = True) -> Dict:
    """`steps` timed config-1 forwards after `warmup` untimed ones (each "step" of the reference arm is this bounded sample)."""
    r = CpuReference(cfg)
    for _ in range(warmup):
        r.forward_seconds()
    times = [r.forward_seconds() for _ in range(steps)]
    fwd = median(times)
    step_s, ratio = extrapolate_step_seconds(fg, cfg, fwd, r.tokens, tokens_full, text_len)
    out = {
        "kind": "reference", "cores": r.cores, "unit": "steps/s", "value": 1.0 / step_s,
        "sample": (f"UNMODIFIED reference model_fn_wan_video (baseline/_ref, torch CPU fp32, {r.cores} threads) at BASELINE config 1: "
                   f"TI2V-5B random-init, 1 forward, latent 1x48x5x16x16 (S={r.tokens}), {CONFIG1['text_len']} text tokens, no CFG; "
                   f"median of {steps} forwards after {warmup} warm-up = {fwd:.3f} s; headline `value` = the faster (for the CPU) of "
                   f"two extrapolations of the same reference code: that forward scaled by counted FLOPs (x{ratio:.1f}) to S={tokens_full}, "
                   f"or one reference DiTBlock timed at 4096 tokens scaled ~S / ~S^2; x2 forwards per CFG step"),
        "measured": {"config": "BASELINE.json configs[0]", "seconds_per_forward": fwd, "forwards_per_s": 1.0 / fwd,
                     "all_forward_seconds": [round(t, 4) for t in times], "tokens": r.tokens, "dtype": "f32",
                     "model_build_s": round(r.build_s, 2)},
        "extrapolated": {"to": f"S={tokens_full}, CFG on (2 forwards per step)", "by": "counted FLOPs (SURVEY 8d)", "flop_ratio": ratio,
                         "seconds_per_step": step_s},
    }
    if with_block_sample:
        s, tot, att = r.block_sample()
        k = tokens_full / s
        blk = (tot - att) * k + att * k * k
        blk_step = blk * cfg.num_layers * 2
        out["extrapolated"]["by_block_sample"] = {
            "what": f"one reference DiTBlock at {s} tokens: {tot:.3f} s of which self-attention {att:.3f} s; rest scaled x{k:.2f}, "
                    f"attention x{k * k:.1f}, x{cfg.num_layers} blocks x2 forwards",
            "seconds_per_step": blk_step}
        # config 1 streams 20 GB of fp32 weights for 320 tokens (memory-bound), so scaling it by FLOPs can OVER-state the CPU's
        # headline time; the line's `value` takes whichever extrapolation is kinder to the CPU
        if blk_step < step_s:
            out["value"] = 1.0 / blk_step
            out["extrapolated"]["seconds_per_step"] = blk_step
            out["extrapolated"]["by"] = "reference DiTBlock sample at 4096 tokens (faster for the CPU than the FLOP-scaled config-1 forward)"
            out["extrapolated"]["by_counted_flops_seconds_per_step"] = step_s
    return out


# ------------------------------------------------------------------------------------------------
# GPU: the library kernels the reference lowers to, and the reference's own forward, on the same B200
# ------------------------------------------------------------------------------------------------
def _time_cuda(fn, iters: int, warmup: int = 3) -> float:
    """Median milliseconds of fn() over `iters` runs, CUDA events on the current stream."""
    import torch

    for _ in range(warmup):
        fn()
    evs = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return median(a.elapsed_time(b) for a, b in evs)


def _time_sustained(fns: Dict, seconds: float = 0.3, rounds: int = 2) -> Dict[str, float]:
    """ms per call of every fn, each run back-to-back for ~`seconds` (so the 1 kW power cap settles the clock the way it does
    inside a denoise step — a handful of isolated calls runs at burst clocks or not, depending on what ran just before), the
    candidates taking turns `rounds` times so that none of them always runs on the hotter chip; mean over the rounds."""
    import torch

    acc = {name: [] for name in fns}
    for _ in range(rounds):
        for name, fn in fns.items():
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            n = max(4, min(400, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3))))
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            acc[name].append(e0.elapsed_time(e1) / n)
    return {name: sum(v) / len(v) for name, v in acc.items()}


def gemm_vs_cublas(fg, cfg, tokens: int, dev, iters: int = 8) -> Dict:
    """The five GEMM shapes of one DiT block at `tokens` rows: ours (fgb_gemm_bf16, bias epilogue) vs ``F.linear`` (cuBLASLt, bias
    epilogue) on the same operands, sustained and alternating (_time_sustained).  Operands of one shape exceed the 126 MB L2."""
    import torch
    import torch.nn.functional as F

    from fairygen_b200 import ops

    d, f = cfg.dim, cfg.ffn_dim
    shapes = [("qkv", 3 * d, d), ("attn_o / cross_q / cross_o", d, d), ("ffn1", f, d), ("ffn2", d, f)]
    out = {}
    g = torch.Generator(device=dev).manual_seed(11)
    for name, n, k in shapes:
        a = torch.randn(tokens, k, device=dev, dtype=torch.bfloat16, generator=g)
        w = torch.randn(n, k, device=dev, dtype=torch.bfloat16, generator=g) * (k ** -0.5)
        b = torch.randn(n, device=dev, dtype=torch.bfloat16, generator=g)
        c = torch.empty(tokens, n, device=dev, dtype=torch.bfloat16)
        t = _time_sustained({"ours": lambda: ops.gemm(a, w, b, c), "lib": lambda: F.linear(a, w, b)})
        ours, lib = t["ours"], t["lib"]
        ref_out = F.linear(a, w, b)
        err = float((c.float() - ref_out.float()).norm() / ref_out.float().norm())
        fl = 2.0 * tokens * n * k
        out[f"{name} [{tokens}x{n}x{k}]"] = {"ours_ms": round(ours, 4), "cublas_ms": round(lib, 4), "ours_tflops": round(fl / ours / 1e9, 1),
                                             "cublas_tflops": round(fl / lib / 1e9, 1), "ours_over_cublas": round(lib / ours, 4),
                                             "rel_l2_vs_cublas": err}
        del a, w, b, c, ref_out
    return out


def attn_vs_library(fg, cfg, tokens: int, text_len: int, dev, iters: int = 5) -> Dict:
    """Self-attention (S x S) and cross-attention (S x text_len) of one block: ours (fgb_head_norm_max + fgb_attn_fwd_bounded, the
    engine's default) vs the reference's own ``flash_attention`` (DIT:27-60) on its flash-attn 2 branch (if it launches on this
    GPU) and on its SDPA branch, plus SDPA pinned to the cuDNN and the flash backends."""
    import torch
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel

    from fairygen_b200 import ops

    from . import ref_loader as rl

    ref = rl.load()
    wd = ref.wd
    H, d = cfg.num_heads, cfg.dim
    g = torch.Generator(device=dev).manual_seed(12)
    out = {}
    fa2_ok = rl.select_attention_backend(dev) == "flash_attn_2"
    for name, s_kv in (("self", tokens), ("cross", text_len)):
        q = torch.randn(tokens, d, device=dev, dtype=torch.bfloat16, generator=g)
        k = torch.randn(s_kv, d, device=dev, dtype=torch.bfloat16, generator=g)
        v = torch.randn(s_kv, d, device=dev, dtype=torch.bfloat16, generator=g)
        o = torch.empty(tokens, d, device=dev, dtype=torch.bfloat16)
        kmax2 = torch.zeros(H, device=dev, dtype=torch.float32)

        qmax2 = torch.zeros(H, device=dev, dtype=torch.float32)

        def ours_with_bound_pass():   # a caller without the fused q/k norm pass computes the key bound separately (per-row query bounds)
            ops.head_norm_max(k, kmax2, H)
            ops.attention(q, k, v, o, H, kmax2=kmax2)

        def ours():   # as the engine runs it: the key bound is a by-product of fgb_qk_norm_rope (self) / cached with the context K|V (cross)
            ops.attention(q, k, v, o, H, kmax2=kmax2)

        def ours_head_bound():   # fgb_attn_fwd_bounded_qk with the head-level query bound given (engine: FGB_QMAX, off by default)
            ops.attention(q, k, v, o, H, kmax2=kmax2, qmax2=qmax2)

        ops.head_norm_max(k, kmax2, H)
        ops.head_norm_max(q, qmax2, H)
        q3, k3, v3 = q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0)
        qh, kh, vh = (t.view(-1, H, 128).transpose(0, 1).unsqueeze(0) for t in (q, k, v))
        saved = wd.FLASH_ATTN_2_AVAILABLE

        def ref_fa2():
            wd.FLASH_ATTN_2_AVAILABLE = True
            return wd.flash_attention(q3, k3, v3, H)

        def ref_sdpa():
            wd.FLASH_ATTN_2_AVAILABLE = False
            return wd.flash_attention(q3, k3, v3, H)

        def sdpa_with(backend):
            def run():
                with sdpa_kernel([backend]):
                    return F.scaled_dot_product_attention(qh, kh, vh)
            return run

        cands = {"ours_ms": ours, "ours_with_bound_pass_ms": ours_with_bound_pass, "ours_head_bound_ms": ours_head_bound}
        if fa2_ok:
            cands["reference_flash_attention[flash_attn_2]_ms"] = ref_fa2
        cands["reference_flash_attention[sdpa]_ms"] = ref_sdpa
        libs = {}
        for label, backend in (("sdpa_cudnn_ms", SDPBackend.CUDNN_ATTENTION), ("sdpa_flash_ms", SDPBackend.FLASH_ATTENTION)):
            try:
                sdpa_with(backend)()
                cands[label] = sdpa_with(backend)
            except Exception as e:   # backend not available for this shape / build
                libs[label] = None
                libs[label + "_error"] = str(e).split("\n")[0][:120]
        try:
            timed = _time_sustained(cands)
            ref_o = ref_sdpa()[0]
        finally:
            wd.FLASH_ATTN_2_AVAILABLE = saved
        res = {"ours_ms": round(timed.pop("ours_ms"), 4), "ours_with_bound_pass_ms": round(timed.pop("ours_with_bound_pass_ms"), 4),
               "ours_head_bound_ms": round(timed.pop("ours_head_bound_ms"), 4),
               "timing": "each candidate back-to-back for ~0.3 s, candidates alternating twice (sustained clocks)"}
        libs.update(timed)
        ours()
        res["rel_l2_vs_library"] = float((o.float() - ref_o.float()).norm() / ref_o.float().norm())
        best = min(v for k_, v in libs.items() if k_.endswith("_ms") and v is not None)
        fl = 4.0 * tokens * s_kv * d
        res.update({k_: (round(v, 4) if isinstance(v, float) else v) for k_, v in libs.items()})
        res.update({"best_library_ms": round(best, 4), "ours_over_best_library": round(best / res["ours_ms"], 4),
                    "ours_tflops": round(fl / res["ours_ms"] / 1e9, 1), "best_library_tflops": round(fl / best / 1e9, 1)})
        out[f"{name} [S_q={tokens}, S_kv={s_kv}, {H} heads x 128]"] = res
        del q, k, v, o, q3, k3, v3, qh, kh, vh, ref_o, cands
    return out


def gpu_reference_forward(cfg, state_dict, latents, timestep, context, dev, forwards: int = 2):
    """The unmodified reference forward (model_fn_wan_video, bf16, reference WanModel holding `state_dict`) on `dev`.
    Returns (output, median wall ms per forward incl. its host-side RoPE table build + H2D copy, attention backend)."""
    import torch

    from . import ref_loader as rl

    ref = rl.load()
    backend = rl.select_attention_backend(dev)
    dit = rl.build_wan_model(cfg, state_dict=state_dict)
    times = []
    out = None
    with torch.no_grad():
        for i in range(forwards + 1):
            _sync(dev)
            t0 = time.perf_counter()
            out = ref.wv.model_fn_wan_video(dit=dit, latents=latents, timestep=timestep, context=context, fuse_vae_embedding_in_latents=True)
            _sync(dev)
            if i > 0:
                times.append((time.perf_counter() - t0) * 1e3)
    del dit
    return out, (median(times) if times else None), backend


def gpu_reference_denoise(cfg, state_dict, latents, z0, ctx_pos, ctx_neg, dev, steps, num_inference_steps=50, cfg_scale=5.0, shift=5.0):
    """`steps` iterations of the reference's denoise loop body (PIPE:285-309) with its own scheduler and model_fn, bf16, on `dev`."""
    import torch

    from . import ref_loader as rl

    ref = rl.load()
    rl.select_attention_backend(dev)
    dit = rl.build_wan_model(cfg, state_dict=state_dict)
    sched = ref.fm.FlowMatchScheduler("Wan")
    sched.set_timesteps(num_inference_steps, denoising_strength=1.0, shift=shift)
    lat = latents.clone()
    lat[:, :, 0:1] = z0
    with torch.no_grad():
        for i in steps:
            ts = sched.timesteps[i].unsqueeze(0).to(dtype=torch.bfloat16, device=dev)
            npos = ref.wv.model_fn_wan_video(dit=dit, latents=lat, timestep=ts, context=ctx_pos, fuse_vae_embedding_in_latents=True)
            nneg = ref.wv.model_fn_wan_video(dit=dit, latents=lat, timestep=ts, context=ctx_neg, fuse_vae_embedding_in_latents=True)
            pred = nneg + cfg_scale * (npos - nneg)
            lat = sched.step(pred, sched.timesteps[i], lat)
            lat[:, :, 0:1] = z0
    del dit
    return lat


def gpu_reference_block(fg, cfg, engine, state_dict, latents, timestep, context, dev, tokens: int, text_len: int,
                        our_steps_per_s: Optional[float]) -> Dict:
    """Everything the N = 1 bench line reports about the reference on the same GPU."""
    import torch

    from . import ref_loader as rl

    if not rl.available():
        return {"unavailable": "baseline/_ref is empty (python baseline/install_ref.py in the build container)"}
    block: Dict = {}
    ref_out, fwd_ms, backend = gpu_reference_forward(cfg, state_dict, latents, timestep, context, dev)
    ours = engine.forward(latents, timestep, context, True)
    rel = float((ours.float() - ref_out.float()).norm() / ref_out.float().norm())
    block.update({
        "what": "UNMODIFIED reference model_fn_wan_video + WanModel (baseline/_ref) in bf16 on this B200, same weights and inputs",
        "attention_backend": backend, "forward_ms": round(fwd_ms, 2), "steps_per_s": 1e3 / (2 * fwd_ms),
        "rel_l2_ours_vs_reference_bf16": rel, "tolerance": 1e-2, "shape": list(latents.shape),
    })
    if our_steps_per_s:
        block["ours_over_reference"] = round(our_steps_per_s / block["steps_per_s"], 3)
    del ref_out, ours
    torch.cuda.empty_cache()
    block["gemm_vs_cublas"] = gemm_vs_cublas(fg, cfg, tokens, dev)
    block["attn_vs_library"] = attn_vs_library(fg, cfg, tokens, text_len, dev)
    return block
