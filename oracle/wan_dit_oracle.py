"""CPU/PyTorch ORACLE for the Wan2.2-TI2V-5B DiT denoise hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch ops over a flat ``{name: tensor}`` weight dict)
of the algorithm the reference runs for this path.  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker / CPU baseline.  The product (``fairygen_b200``) never imports
it and has no CPU fallback.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference from
``/root/reference/animation`` (in the build container) and stores its outputs for seeded inputs
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function here against those
vectors.  Third-party pieces that are not in the reference tree (``xfuser`` for Ulysses SP —
unpinned, absent) are restated from the reference's call sites and noted below.

Reference files restated (relative to /root/reference/animation/diffsynth):
  DIT  = models/wan_video_dit.py        PIPE = pipelines/wan_video.py
  FM   = diffusion/flow_match.py        LORA = utils/lora/general.py
  USP  = utils/xfuser/xdit_context_parallel.py
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------
# configuration                                                    configs/model_configs.py:290-295
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class DiTConfig:
    dim: int = 3072
    in_dim: int = 48
    ffn_dim: int = 14336
    out_dim: int = 48
    text_dim: int = 4096
    freq_dim: int = 256
    eps: float = 1e-6
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    num_heads: int = 24
    num_layers: int = 30
    seperated_timestep: bool = True  # (sic) the reference's spelling

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads


TI2V_5B = DiTConfig()
TINY = DiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)


def param_shapes(cfg: DiTConfig) -> Dict[str, Tuple[int, ...]]:
    """State-dict keys and shapes of the reference WanModel (DIT:271-336) for a TI2V config."""
    d, f = cfg.dim, cfg.ffn_dim
    pt, ph, pw = cfg.patch_size
    s: Dict[str, Tuple[int, ...]] = {
        "patch_embedding.weight": (d, cfg.in_dim, pt, ph, pw),
        "patch_embedding.bias": (d,),
        "text_embedding.0.weight": (d, cfg.text_dim),
        "text_embedding.0.bias": (d,),
        "text_embedding.2.weight": (d, d),
        "text_embedding.2.bias": (d,),
        "time_embedding.0.weight": (d, cfg.freq_dim),
        "time_embedding.0.bias": (d,),
        "time_embedding.2.weight": (d, d),
        "time_embedding.2.bias": (d,),
        "time_projection.1.weight": (6 * d, d),
        "time_projection.1.bias": (6 * d,),
        "head.head.weight": (cfg.out_dim * pt * ph * pw, d),
        "head.head.bias": (cfg.out_dim * pt * ph * pw,),
        "head.modulation": (1, 2, d),
    }
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        for attn in ("self_attn", "cross_attn"):
            for proj in ("q", "k", "v", "o"):
                s[f"{b}{attn}.{proj}.weight"] = (d, d)
                s[f"{b}{attn}.{proj}.bias"] = (d,)
            s[f"{b}{attn}.norm_q.weight"] = (d,)
            s[f"{b}{attn}.norm_k.weight"] = (d,)
        s[b + "norm3.weight"] = (d,)
        s[b + "norm3.bias"] = (d,)
        s[b + "ffn.0.weight"] = (f, d)
        s[b + "ffn.0.bias"] = (f,)
        s[b + "ffn.2.weight"] = (d, f)
        s[b + "ffn.2.bias"] = (d,)
        s[b + "modulation"] = (1, 6, d)
    return s


def _seed_for(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF


def make_weights(cfg: DiTConfig, seed: int = 0, dtype=torch.float32, device="cpu") -> Weights:
    """Deterministic synthetic weights, one generator per tensor keyed by its NAME, so the same
    tensors can be regenerated anywhere without shipping them.  Distributions follow PyTorch's
    default init of the reference modules (uniform(+-1/sqrt(fan_in)) for Linear/Conv, randn/sqrt(dim)
    for modulation, DIT:210,259) except that norm weights/biases are perturbed off 1/0 so a wrong
    or missing affine term cannot pass the parity tests."""
    out: Weights = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator(device=device).manual_seed(_seed_for(name, seed))
        if name.endswith("modulation"):
            w = torch.randn(shape, generator=g, device=device) / math.sqrt(cfg.dim)
        elif "norm" in name:
            base = 1.0 if name.endswith("weight") else 0.0
            w = base + 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            fan_in = math.prod(shape[1:]) if len(shape) > 1 else None
            if fan_in is None:  # bias: fan_in of the matching weight
                wshape = param_shapes(cfg)[name[: -len("bias")] + "weight"]
                fan_in = math.prod(wshape[1:])
            bound = 1.0 / math.sqrt(fan_in)
            w = (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound
        out[name] = w.to(dtype)
    return out


def make_lora(cfg: DiTConfig, rank: int = 32, seed: int = 2, dtype=torch.float32, device="cpu") -> Weights:
    """Synthetic motion LoRA in the on-disk key format the reference loads (LORA:10-41):
    ``blocks.N.{self_attn,cross_attn}.{q,k,v,o}.lora_{A,B}.default.weight`` and ``ffn.{0,2}``."""
    shapes = param_shapes(cfg)
    out: Weights = {}
    for i in range(cfg.num_layers):
        targets = [f"blocks.{i}.{a}.{p}" for a in ("self_attn", "cross_attn") for p in "qkvo"]
        targets += [f"blocks.{i}.ffn.0", f"blocks.{i}.ffn.2"]
        for t in targets:
            n, k = shapes[t + ".weight"]
            ga = torch.Generator(device=device).manual_seed(_seed_for(t + ".A", seed))
            gb = torch.Generator(device=device).manual_seed(_seed_for(t + ".B", seed))
            out[f"{t}.lora_A.default.weight"] = (torch.randn((rank, k), generator=ga, device=device) / math.sqrt(k)).to(dtype)
            out[f"{t}.lora_B.default.weight"] = (torch.randn((n, rank), generator=gb, device=device) * 0.02).to(dtype)
    return out


# --------------------------------------------------------------------------------------------
# LoRA fuse                                                                        LORA:10-62
# --------------------------------------------------------------------------------------------
def lora_target_names(lora: Weights) -> Dict[str, Tuple[str, str]]:
    """module name -> (B key, A key), following the key normalisation of LORA:10-30."""
    names: Dict[str, Tuple[str, str]] = {}
    for key in lora:
        a_tag, b_tag = ("lora_down", "lora_up") if ".lora_up." in key else ("lora_A", "lora_B")
        if b_tag not in key:
            continue
        parts = key.split(".")
        at = parts.index(b_tag)
        if len(parts) > at + 2:  # drop the adapter name ("default")
            parts.pop(at + 1)
        parts.pop(at)
        if parts[0] == "diffusion_model":
            parts.pop(0)
        parts.pop(-1)  # "weight"
        names[".".join(parts)] = (key, key.replace(b_tag, a_tag))
    return names


def fuse_lora(weights: Weights, lora: Weights, alpha: float = 1.0) -> Weights:
    """W <- W + alpha * (B @ A), computed in the dtype of W as the reference does (LORA:44-62)."""
    out = dict(weights)
    for module, (kb, ka) in lora_target_names(lora).items():
        wkey = module + ".weight"
        if wkey not in out:
            continue
        w = out[wkey]
        up = lora[kb].to(device=w.device, dtype=w.dtype)
        down = lora[ka].to(device=w.device, dtype=w.dtype)
        out[wkey] = w + alpha * torch.mm(up, down)
    return out


# --------------------------------------------------------------------------------------------
# primitives                                                                      DIT:63-110
# --------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """[cos(p w_i), sin(p w_i)] with w_i = 10000^(-i/(dim/2)), evaluated in fp64 (DIT:67-71)."""
    half = dim // 2
    w = torch.pow(10000, -torch.arange(half, dtype=torch.float64, device=position.device).div(half))
    ang = torch.outer(position.to(torch.float64), w)
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1).to(position.dtype)


def rope_axis_table(dim: int, end: int = 1024, theta: float = 10000.0) -> torch.Tensor:
    """complex128 [end, dim/2] unit phasors for one axis (DIT:82-88)."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].double() / dim))
    ang = torch.outer(torch.arange(end), inv)
    return torch.polar(torch.ones_like(ang), ang)


def rope_tables_3d(head_dim: int):
    """Axis split of the head_dim/2 complex lanes: frame gets dim-2*(dim//3), row/col dim//3 (DIT:74-79)."""
    return (rope_axis_table(head_dim - 2 * (head_dim // 3)), rope_axis_table(head_dim // 3), rope_axis_table(head_dim // 3))


def rope_freqs(tables, f: int, h: int, w: int) -> torch.Tensor:
    """(f*h*w, 1, head_dim/2) complex128 in (f h w) token order (PIPE:1271-1275)."""
    tf, th, tw = tables
    return torch.cat(
        [
            tf[:f].view(f, 1, 1, -1).expand(f, h, w, -1),
            th[:h].view(1, h, 1, -1).expand(f, h, w, -1),
            tw[:w].view(1, 1, w, -1).expand(f, h, w, -1),
        ],
        dim=-1,
    ).reshape(f * h * w, 1, -1)


def rope_apply(x: torch.Tensor, freqs: torch.Tensor, num_heads: int) -> torch.Tensor:
    """Adjacent pairs as complex numbers times the phasor, in fp64, back to x.dtype (DIT:91-96)."""
    b, s, _ = x.shape
    xc = torch.view_as_complex(x.to(torch.float64).reshape(b, s, num_heads, -1, 2))
    return torch.view_as_real(xc * freqs.to(x.device)).flatten(2).to(x.dtype)


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """fp32 x*rsqrt(mean(x^2)+eps) over the FULL last dim, cast back, then * weight (DIT:105-110)."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).to(x.dtype) * weight


def layer_norm(x: torch.Tensor, eps: float, weight=None, bias=None) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def modulate(x, shift, scale):
    return x * (1 + scale) + shift  # DIT:63-64


def attention(q, k, v, num_heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v per head, no mask (DIT:54-59 — the in-tree SDPA branch)."""
    b, sq, _ = q.shape
    qh = q.view(b, sq, num_heads, -1).transpose(1, 2)
    kh = k.view(b, k.shape[1], num_heads, -1).transpose(1, 2)
    vh = v.view(b, v.shape[1], num_heads, -1).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh)
    return o.transpose(1, 2).reshape(b, sq, -1)


def linear(w: Weights, name: str, x: torch.Tensor) -> torch.Tensor:
    hook = w.get("__linear__")  # oracle/wan_train_oracle.py installs the stage-2 LoRA forward here
    if hook is not None:
        return hook(w, name, x)
    return F.linear(x, w[name + ".weight"], w[name + ".bias"])


# --------------------------------------------------------------------------------------------
# DiT block, head                                                        DIT:123-229, 252-268
# --------------------------------------------------------------------------------------------
def self_attention(w: Weights, p: str, x, freqs, cfg: DiTConfig, attn_fn=attention):
    q = rms_norm(linear(w, p + "q", x), w[p + "norm_q.weight"], cfg.eps)
    k = rms_norm(linear(w, p + "k", x), w[p + "norm_k.weight"], cfg.eps)
    v = linear(w, p + "v", x)
    q = rope_apply(q, freqs, cfg.num_heads)
    k = rope_apply(k, freqs, cfg.num_heads)
    return linear(w, p + "o", attn_fn(q, k, v, cfg.num_heads))


def cross_attention(w: Weights, p: str, x, ctx, cfg: DiTConfig):
    q = rms_norm(linear(w, p + "q", x), w[p + "norm_q.weight"], cfg.eps)
    k = rms_norm(linear(w, p + "k", ctx), w[p + "norm_k.weight"], cfg.eps)
    v = linear(w, p + "v", ctx)
    return linear(w, p + "o", attention(q, k, v, cfg.num_heads))


def dit_block(w: Weights, i: int, x, context, t_mod, freqs, cfg: DiTConfig, attn_fn=attention):
    """DIT:213-229.  t_mod is (B,6,D) or per-token (B,S,6,D)."""
    p = f"blocks.{i}."
    per_token = t_mod.dim() == 4
    mods = (w[p + "modulation"].to(dtype=t_mod.dtype, device=t_mod.device) + t_mod).chunk(6, dim=2 if per_token else 1)
    if per_token:
        mods = [m.squeeze(2) for m in mods]
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mods
    a = modulate(layer_norm(x, cfg.eps), shift_msa, scale_msa)
    x = x + gate_msa * self_attention(w, p + "self_attn.", a, freqs, cfg, attn_fn)
    x = x + cross_attention(w, p + "cross_attn.", layer_norm(x, cfg.eps, w[p + "norm3.weight"], w[p + "norm3.bias"]), context, cfg)
    a = modulate(layer_norm(x, cfg.eps), shift_mlp, scale_mlp)
    hidden = F.gelu(linear(w, p + "ffn.0", a), approximate="tanh")
    return x + gate_mlp * linear(w, p + "ffn.2", hidden)


def head(w: Weights, x, t, cfg: DiTConfig):
    """DIT:261-268; t is (B,S,D) per-token or (B,D)."""
    mod = w["head.modulation"].to(dtype=t.dtype, device=t.device)
    if t.dim() == 3:
        shift, scale = (mod.unsqueeze(0) + t.unsqueeze(2)).chunk(2, dim=2)
        y = layer_norm(x, cfg.eps) * (1 + scale.squeeze(2)) + shift.squeeze(2)
    else:
        shift, scale = (mod + t.unsqueeze(1)).chunk(2, dim=1)
        y = layer_norm(x, cfg.eps) * (1 + scale) + shift
    return linear(w, "head.head", y)


def unpatchify(x, grid, cfg: DiTConfig):
    """'b (f h w) (x y z c) -> b c (f x) (h y) (w z)'  (DIT:346-351)."""
    f, h, wd = grid
    pt, ph, pw = cfg.patch_size
    b = x.shape[0]
    x = x.view(b, f, h, wd, pt, ph, pw, cfg.out_dim)
    return x.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(b, cfg.out_dim, f * pt, h * ph, wd * pw)


# --------------------------------------------------------------------------------------------
# one DiT forward = model_fn_wan_video, TI2V branches                    PIPE:1217-1388
# --------------------------------------------------------------------------------------------
def time_embeddings(w: Weights, cfg: DiTConfig, latents, timestep, fuse_vae_embedding_in_latents: bool):
    """Returns (t, t_mod).  Per-token (first latent frame at t=0) when the TI2V flag is set (PIPE:1218-1231)."""
    if cfg.seperated_timestep and fuse_vae_embedding_in_latents:
        hw4 = latents.shape[3] * latents.shape[4] // 4
        ts = torch.concat(
            [
                torch.zeros((1, hw4), dtype=latents.dtype, device=latents.device),
                torch.ones((latents.shape[2] - 1, hw4), dtype=latents.dtype, device=latents.device) * timestep,
            ]
        ).flatten()
        e = sinusoidal_embedding_1d(cfg.freq_dim, ts).unsqueeze(0)
        t = linear(w, "time_embedding.2", F.silu(linear(w, "time_embedding.0", e)))
        t_mod = linear(w, "time_projection.1", F.silu(t)).unflatten(2, (6, cfg.dim))
    else:
        e = sinusoidal_embedding_1d(cfg.freq_dim, timestep)
        t = linear(w, "time_embedding.2", F.silu(linear(w, "time_embedding.0", e)))
        t_mod = linear(w, "time_projection.1", F.silu(t)).unflatten(1, (6, cfg.dim))
    return t, t_mod


def text_embedding(w: Weights, context):
    return linear(w, "text_embedding.2", F.gelu(linear(w, "text_embedding.0", context), approximate="tanh"))


def patch_embed(w: Weights, latents, cfg: DiTConfig):
    x = F.conv3d(latents, w["patch_embedding.weight"], w["patch_embedding.bias"], stride=cfg.patch_size)
    f, h, wd = x.shape[2:]
    return x.flatten(2).transpose(1, 2).contiguous(), (f, h, wd)  # 'b c f h w -> b (f h w) c'


def dit_forward(
    w: Weights,
    cfg: DiTConfig,
    latents: torch.Tensor,
    timestep: torch.Tensor,
    context: torch.Tensor,
    fuse_vae_embedding_in_latents: bool = True,
    rope_tables=None,
    return_hidden: bool = False,
    attn_fn=attention,
) -> torch.Tensor:
    """latents (1,C,F,H,W), timestep (1,), context (1,L,text_dim) -> velocity prediction like latents."""
    t, t_mod = time_embeddings(w, cfg, latents, timestep, fuse_vae_embedding_in_latents)
    ctx = text_embedding(w, context)
    x, (f, h, wd) = patch_embed(w, latents, cfg)
    tables = rope_tables if rope_tables is not None else rope_tables_3d(cfg.head_dim)
    freqs = rope_freqs(tables, f, h, wd).to(x.device)
    for i in range(cfg.num_layers):
        x = dit_block(w, i, x, ctx, t_mod, freqs, cfg, attn_fn)
    if return_hidden:
        return x
    x = head(w, x, t, cfg)
    return unpatchify(x, (f, h, wd), cfg)


# --------------------------------------------------------------------------------------------
# Ulysses sequence parallel, emulated with "virtual ranks" in one process
#   reference glue: PIPE:1224-1227, 1310-1315, 1379-1382; USP:30-55, 125-146.  The all-to-all
#   itself lives in xfuser (NOT in the reference tree, unpinned): restated from its call site as
#   "scatter heads / gather sequence, attend over the full sequence, inverse".
# --------------------------------------------------------------------------------------------
def sp_chunk_pad(x: torch.Tensor, world: int, dim: int = 1):
    """torch.chunk + zero-pad every chunk to the first chunk's length (PIPE:1312-1315). Returns (chunks, pad)."""
    chunks = list(torch.chunk(x, world, dim=dim))
    n0 = chunks[0].shape[dim]
    pad = n0 - chunks[-1].shape[dim]
    out = []
    for c in chunks:
        extra = n0 - c.shape[dim]
        if extra:
            shape = list(c.shape)
            shape[dim] = extra
            c = torch.cat([c, c.new_zeros(shape)], dim=dim)
        out.append(c)
    while len(out) < world:  # torch.chunk may return fewer chunks
        out.append(torch.zeros_like(out[0]))
    return out, pad


def sp_rope_apply(x, freqs, num_heads, rank: int, world: int):
    """USP:42-55 — the rank's slice of a ones-padded phasor table."""
    s_local = x.shape[1]
    total = s_local * world
    if freqs.shape[0] < total:
        freqs = torch.cat([freqs, torch.ones(total - freqs.shape[0], *freqs.shape[1:], dtype=freqs.dtype)], dim=0)
    return rope_apply(x, freqs[rank * s_local : (rank + 1) * s_local], num_heads)


def ulysses_attention(qs, ks, vs, num_heads: int, valid_tokens: Optional[int] = None):
    """qs/ks/vs: per-rank lists of (1, s_local, D).  Head-scatter/sequence-gather all-to-all, local
    attention over the full sequence on heads/world heads, inverse all-to-all.  ``valid_tokens``
    masks the padded keys (our native behaviour); None reproduces the reference, which lets the
    zero-padded rows attend."""
    world = len(qs)
    d = qs[0].shape[-1] // num_heads
    hpr = num_heads // world
    outs = [[None] * world for _ in range(world)]
    for r in range(world):  # rank r owns heads [r*hpr, (r+1)*hpr) over the whole (padded) sequence
        sl = slice(r * hpr * d, (r + 1) * hpr * d)
        q = torch.cat([t[..., sl] for t in qs], dim=1)
        k = torch.cat([t[..., sl] for t in ks], dim=1)
        v = torch.cat([t[..., sl] for t in vs], dim=1)
        if valid_tokens is not None:
            k, v = k[:, :valid_tokens], v[:, :valid_tokens]
        o = attention(q, k, v, hpr)
        for src, piece in enumerate(torch.chunk(o, world, dim=1)):
            outs[src][r] = piece
    return [torch.cat(outs[r], dim=-1) for r in range(world)]


# --------------------------------------------------------------------------------------------
# flow-match scheduler + denoise loop                         FM:29-39, 132-154; PIPE:282-309
# --------------------------------------------------------------------------------------------
def flow_match_schedule(num_inference_steps: int = 50, denoising_strength: float = 1.0, shift: float = 5.0):
    """'Wan' template: sigma = linspace(start, 0, n+1)[:-1], shifted; t = 1000 sigma (FM:29-39)."""
    sigma_start = 0.0 + (1.0 - 0.0) * denoising_strength
    sigmas = torch.linspace(sigma_start, 0.0, num_inference_steps + 1)[:-1]
    sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
    return sigmas, sigmas * 1000


def flow_match_step(model_output, step_index: int, sample, sigmas):
    """x + v * (sigma_next - sigma); sigma_next = 0 after the last step (FM:144-154)."""
    sigma = sigmas[step_index]
    sigma_next = sigmas[step_index + 1] if step_index + 1 < len(sigmas) else 0
    return sample + model_output * (sigma_next - sigma)


def cfg_combine(noise_pos, noise_neg, cfg_scale: float):
    return noise_neg + cfg_scale * (noise_pos - noise_neg)  # PIPE:302


def denoise(
    w: Weights,
    cfg: DiTConfig,
    latents,
    context_pos,
    context_neg,
    first_frame_latents=None,
    num_inference_steps: int = 50,
    cfg_scale: float = 5.0,
    shift: float = 5.0,
    forward=None,
):
    """The hot loop of WanVideoPipeline.__call__ (PIPE:285-309) for TI2V."""
    fwd = forward if forward is not None else (lambda lat, ts, ctx: dit_forward(w, cfg, lat, ts, ctx, first_frame_latents is not None))
    sigmas, timesteps = flow_match_schedule(num_inference_steps, 1.0, shift)
    for i, ts in enumerate(timesteps):
        ts_in = ts.unsqueeze(0).to(dtype=latents.dtype, device=latents.device)  # PIPE:293 (bf16 rounding of t)
        npos = fwd(latents, ts_in, context_pos)
        if cfg_scale != 1.0:
            nneg = fwd(latents, ts_in, context_neg)
            npred = cfg_combine(npos, nneg, cfg_scale)
        else:
            npred = npos
        latents = flow_match_step(npred, i, latents, sigmas)
        if first_frame_latents is not None:
            latents[:, :, 0:1] = first_frame_latents
    return latents


# --------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------
def latent_shape(cfg: DiTConfig, height: int, width: int, num_frames: int):
    return (1, cfg.in_dim, (num_frames - 1) // 4 + 1, height // 16, width // 16)


def make_inputs(cfg: DiTConfig, shape, dtype=torch.float32, text_len: int = 512, live_text: int = 64):
    """latents seed 1, first-frame latents seed 3, positive/negative context seeds 4/5 with rows >=
    live_text zeroed (mirrors the prompt-length zeroing of PIPE:410-411)."""
    def rn(s, seed):
        return torch.randn(s, generator=torch.Generator("cpu").manual_seed(seed), dtype=torch.float32)

    lat = rn(shape, 1).to(dtype)
    z0 = rn((shape[0], shape[1], 1, shape[3], shape[4]), 3).to(dtype)
    cp = rn((1, text_len, cfg.text_dim), 4)
    cn = rn((1, text_len, cfg.text_dim), 5)
    cp[:, live_text:] = 0
    cn[:, live_text:] = 0
    return lat, z0, cp.to(dtype), cn.to(dtype)


def counted_flops(cfg: DiTConfig, s: int, text_len: int = 512) -> float:
    """Algorithmic FLOPs of one DiT forward: x-GEMMs, self/cross attention, patch+head."""
    d, f = cfg.dim, cfg.ffn_dim
    per_block = 12 * s * d * d + 4 * s * d * f + 4 * s * s * d + 4 * s * text_len * d
    io = cfg.out_dim * math.prod(cfg.patch_size)
    return cfg.num_layers * per_block + 2 * (2 * s * d * io)
