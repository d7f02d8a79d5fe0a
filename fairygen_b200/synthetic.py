"""Synthetic weights and inputs of the reference's shapes (there are no checkpoints or datasets on the
GPU box).  State-dict names/shapes are those of the reference ``WanModel`` (animation/diffsynth/models/
wan_video_dit.py:271-336); values follow PyTorch's default init of those modules (uniform +-1/sqrt(fan_in)
for Linear/Conv3d, randn/sqrt(dim) for the modulation tables, DIT:210,259), with norm affine terms
perturbed off (1, 0) so they cannot be skipped unnoticed.  "Merged motion LoRA" = rank-32 B@A added to
the 300 target Linears exactly as ``GeneralLoRALoader.fuse_lora_to_base_model`` does (utils/lora/general.py:44-62).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import torch

from .config import WanDiTConfig


def param_shapes(cfg: WanDiTConfig) -> Dict[str, Tuple[int, ...]]:
    d, f = cfg.dim, cfg.ffn_dim
    pt, ph, pw = cfg.patch_size
    io = cfg.out_dim * pt * ph * pw
    s: Dict[str, Tuple[int, ...]] = {
        "patch_embedding.weight": (d, cfg.in_dim, pt, ph, pw), "patch_embedding.bias": (d,),
        "text_embedding.0.weight": (d, cfg.text_dim), "text_embedding.0.bias": (d,),
        "text_embedding.2.weight": (d, d), "text_embedding.2.bias": (d,),
        "time_embedding.0.weight": (d, cfg.freq_dim), "time_embedding.0.bias": (d,),
        "time_embedding.2.weight": (d, d), "time_embedding.2.bias": (d,),
        "time_projection.1.weight": (6 * d, d), "time_projection.1.bias": (6 * d,),
        "head.head.weight": (io, d), "head.head.bias": (io,), "head.modulation": (1, 2, d),
    }
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        for attn in ("self_attn", "cross_attn"):
            for proj in "qkvo":
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


def lora_targets(cfg: WanDiTConfig):
    """The 10 Linears per block the motion LoRA adapts: "q,k,v,o,ffn.0,ffn.2" (training_module.py:180-185)."""
    for i in range(cfg.num_layers):
        for a in ("self_attn", "cross_attn"):
            for p in "qkvo":
                yield f"blocks.{i}.{a}.{p}"
        yield f"blocks.{i}.ffn.0"
        yield f"blocks.{i}.ffn.2"


def _gen(name: str, seed: int, device) -> torch.Generator:
    return torch.Generator(device=device).manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def random_state_dict(cfg: WanDiTConfig, seed: int = 0, device="cuda", dtype=torch.bfloat16, lora_rank: int = 0,
                      lora_seed: int = 2) -> Dict[str, torch.Tensor]:
    """Random-init weights generated tensor by tensor on `device`; with lora_rank > 0 a synthetic motion
    LoRA is merged into the target weights (W += B @ A in `dtype`)."""
    shapes = param_shapes(cfg)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        g = _gen(name, seed, device)
        if name.endswith("modulation"):
            w = torch.randn(shape, generator=g, device=device) / math.sqrt(cfg.dim)
        elif "norm" in name:
            w = (1.0 if name.endswith("weight") else 0.0) + 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            wshape = shape if len(shape) > 1 else shapes[name[: -len("bias")] + "weight"]
            bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
            w = (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound
        out[name] = w.to(dtype)
    if lora_rank > 0:
        for t in lora_targets(cfg):
            n, k = shapes[t + ".weight"]
            a = (torch.randn((lora_rank, k), generator=_gen(t + ".A", lora_seed, device), device=device) / math.sqrt(k)).to(dtype)
            b = (torch.randn((n, lora_rank), generator=_gen(t + ".B", lora_seed, device), device=device) * 0.02).to(dtype)
            out[t + ".weight"] = out[t + ".weight"] + torch.mm(b, a)
    return out


def random_lora(cfg: WanDiTConfig, rank: int = 32, seed: int = 2, device="cuda", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Unmerged synthetic motion LoRA (stage-1 A1, B1) in the on-disk key format the reference loads
    (``<module>.lora_{A,B}.default.weight``, utils/lora/general.py:10-41) — the same tensors random_state_dict merges."""
    shapes = param_shapes(cfg)
    out: Dict[str, torch.Tensor] = {}
    for t in lora_targets(cfg):
        n, k = shapes[t + ".weight"]
        out[f"{t}.lora_A.default.weight"] = (torch.randn((rank, k), generator=_gen(t + ".A", seed, device), device=device) / math.sqrt(k)).to(dtype)
        out[f"{t}.lora_B.default.weight"] = (torch.randn((n, rank), generator=_gen(t + ".B", seed, device), device=device) * 0.02).to(dtype)
    return out


def latent_shape(cfg: WanDiTConfig, height: int, width: int, num_frames: int):
    """(1, C, (frames-1)/4+1, H/16, W/16) — Wan2.2 VAE38 compression (wan_video_vae.py:1354-1382)."""
    return (1, cfg.in_dim, (num_frames - 1) // 4 + 1, height // 16, width // 16)


def synthetic_inputs(cfg: WanDiTConfig, shape, text_len: int = 512, live_text: int = 64, pin: bool = True):
    """Host tensors (bf16, pinned): latents seed 1, first-frame latents seed 3, positive / negative context
    seeds 4 / 5 with rows >= live_text zeroed (the prompt-length zeroing of wan_video.py:410-411)."""
    def rn(s, seed):
        return torch.randn(s, generator=torch.Generator("cpu").manual_seed(seed), dtype=torch.float32)

    lat = rn(shape, 1).to(torch.bfloat16)
    z0 = rn((shape[0], shape[1], 1, shape[3], shape[4]), 3).to(torch.bfloat16)
    cp, cn = rn((1, text_len, cfg.text_dim), 4), rn((1, text_len, cfg.text_dim), 5)
    cp[:, live_text:] = 0
    cn[:, live_text:] = 0
    outs = [lat, z0, cp.to(torch.bfloat16), cn.to(torch.bfloat16)]
    if pin and torch.cuda.is_available():
        outs = [t.pin_memory() for t in outs]
    return tuple(outs)
