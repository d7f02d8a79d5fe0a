"""CPU/PyTorch ORACLE for the stage-2 motion-LoRA fine-tune step (BASELINE config 5).  TEST INFRASTRUCTURE ONLY.

Restates, on top of ``wan_dit_oracle`` (the pinned DiT forward), what the reference trains:

  * the stage-2 LoRA forward that ``DiffusionTrainingModule`` monkey-patches onto every adapted Linear
    (animation/diffsynth/diffusion/training_module.py:317-352, "TMOD"):
        y = W x + b + s * B1(A1 x) + s * F.linear(A1 x, B2 * mask * 2),   s = alpha / r = 1,
    mask ~ Bernoulli(0.5) drawn fresh per layer per call (TMOD:338-342) — INJECTED here,
    only ``lora_B2.weight`` trainable (TMOD:279-307);
  * ``FlowMatchSFTLoss`` (diffusion/loss.py:5-21, "LOSS") with the scheduler's training helpers
    (diffusion/flow_match.py:132-142, 164-179): noisy latents, target ``noise - x0``, fp32 MSE times the
    timestep weight.  As in the reference the latents are noised AFTER the clean first frame was planted and
    the per-token timestep still marks first-frame tokens t = 0.

Gradients come from torch autograd, exactly as in the reference.  ``peft`` (which supplies the LoRA layer class the
reference patches) is not installed here and unpinned by the reference, so the LoRA arithmetic is restated from
TMOD; ``oracle/make_golden_train.py`` pins it by wrapping the REAL reference ``WanModel``'s Linears with a module
that executes TMOD:317-352 literally and storing the real model's loss / B2 gradients (tests/golden/train.npz).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import wan_dit_oracle as o

Weights = Dict[str, torch.Tensor]


def lora_targets(cfg: o.DiTConfig):
    """The 10 adapted Linears per block: --lora_target_modules "q,k,v,o,ffn.0,ffn.2" (peft suffix match)."""
    for i in range(cfg.num_layers):
        for a in ("self_attn", "cross_attn"):
            for p in "qkvo":
                yield f"blocks.{i}.{a}.{p}"
        yield f"blocks.{i}.ffn.0"
        yield f"blocks.{i}.ffn.2"


def make_b2(cfg: o.DiTConfig, rank: int = 32, seed: int = 6, std: float = 0.02, dtype=torch.float32) -> Weights:
    """Non-zero B2 (the reference zero-initialises it, TMOD:296-298; zero would make every activation-path
    gradient check trivial) — N(0, std^2) per SURVEY §8(d)."""
    shapes = o.param_shapes(cfg)
    out = {}
    for t in lora_targets(cfg):
        g = torch.Generator().manual_seed(o._seed_for(t + ".B2", seed))
        out[t] = (torch.randn((shapes[t + ".weight"][0], rank), generator=g) * std).to(dtype)
    return out


def make_masks(cfg: o.DiTConfig, rank: int = 32, seed: int = 7) -> Dict[str, torch.Tensor]:
    """Fixed keep-masks `rand > 0.5` (uint8), one per adapted Linear."""
    shapes = o.param_shapes(cfg)
    out = {}
    for t in lora_targets(cfg):
        g = torch.Generator().manual_seed(o._seed_for(t + ".mask", seed))
        out[t] = (torch.rand((shapes[t + ".weight"][0], rank), generator=g) > 0.5).to(torch.uint8)
    return out


def stage2_linear_fn(lora: Weights, b2: Weights, masks: Dict[str, torch.Tensor], scaling: float = 1.0, dropout_prob: float = 0.5):
    """Returns the ``__linear__`` hook: TMOD:317-352 for adapted modules, plain nn.Linear otherwise."""
    def hook(w: Weights, name: str, x: torch.Tensor) -> torch.Tensor:
        result = F.linear(x, w[name + ".weight"], w[name + ".bias"])                       # TMOD:320
        if name not in b2:
            return result
        a1 = lora[f"{name}.lora_A.default.weight"].to(x.dtype)
        b1 = lora[f"{name}.lora_B.default.weight"].to(x.dtype)
        result = result + F.linear(F.linear(x, a1), b1) * scaling                         # TMOD:336
        mask = masks[name].to(dtype=b2[name].dtype, device=b2[name].device)               # TMOD:338-342 (injected)
        b2_dropped = b2[name] * mask * (1.0 / (1 - dropout_prob))                         # TMOD:343-344
        update = F.linear(F.linear(x, a1), b2_dropped)                                    # TMOD:346-347
        return result + update * scaling                                                  # TMOD:348
    return hook


def stage1_linear_fn(lora: Weights, masks: Dict[str, torch.Tensor], scaling: float = 1.0, dropout_prob: float = 0.8):
    """The stage-1 ``new_forward`` (TMOD:200-264): y = W x + b + s * F.linear(A x, B * mask * 5), A and B trainable,
    mask ~ Bernoulli(0.2) keep (``rand_like(B) > 0.8``) drawn per layer per call — injected here."""
    def hook(w: Weights, name: str, x: torch.Tensor) -> torch.Tensor:
        result = F.linear(x, w[name + ".weight"], w[name + ".bias"])                       # TMOD:218
        ka, kb = f"{name}.lora_A.default.weight", f"{name}.lora_B.default.weight"
        if kb not in lora:
            return result
        a, b = lora[ka], lora[kb]
        mask = masks[name].to(dtype=b.dtype, device=b.device)                             # TMOD:234-236 (injected)
        b_dropped = b * mask * (1.0 / (1 - dropout_prob))                                 # TMOD:237-238
        update = F.linear(F.linear(x, a), b_dropped)                                      # TMOD:240-241
        return result + update * scaling                                                  # TMOD:242
    return hook


def make_masks_stage1(cfg: o.DiTConfig, rank: int = 32, seed: int = 8) -> Dict[str, torch.Tensor]:
    """Fixed keep-masks `rand > 0.8` (uint8), one per adapted Linear."""
    shapes = o.param_shapes(cfg)
    out = {}
    for t in lora_targets(cfg):
        g = torch.Generator().manual_seed(o._seed_for(t + ".mask1", seed))
        out[t] = (torch.rand((shapes[t + ".weight"][0], rank), generator=g) > 0.8).to(torch.uint8)
    return out


def training_schedule(num_steps: int = 1000, shift: float = 5.0):
    """FlowMatchScheduler('Wan').set_timesteps(1000, training=True) (flow_match.py:29-39, 132-142): sigmas,
    timesteps and the bell-shaped per-timestep loss weights."""
    sigmas = torch.linspace(1.0, 0.0, num_steps + 1)[:-1]
    sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
    timesteps = sigmas * 1000
    steps = 1000
    y = torch.exp(-2 * ((timesteps - steps / 2) / steps) ** 2)
    y = y - y.min()
    weights = y * (steps / y.sum())
    if len(timesteps) != 1000:
        weights = weights * (len(timesteps) / steps)
        weights = weights + weights[1]
    return sigmas, timesteps, weights


def sft_loss(w: Weights, cfg: o.DiTConfig, lora: Weights, b2: Optional[Weights], masks, x0: torch.Tensor, noise: torch.Tensor,
             timestep_id: int, context: torch.Tensor, fuse_vae_embedding_in_latents: bool = True,
             schedule=None, return_pred: bool = False, timestep_dtype=None):
    """LOSS:5-21 with the random draws (timestep id, noise) injected.  x0/noise (1,C,F,H,W) in the compute dtype."""
    sigmas, timesteps, weights = schedule if schedule is not None else training_schedule()
    dtype = x0.dtype
    # LOSS:9 casts the timestep to the PIPELINE dtype (bf16 in training) BEFORE add_noise / training_weight look
    # their index up again by argmin (FM:164-179) — so with bf16 the sigma and the loss weight can belong to a
    # neighbouring schedule entry.  timestep_dtype reproduces that when the oracle itself computes in fp32.
    t_cast = timesteps[timestep_id:timestep_id + 1].to(timestep_dtype or dtype)
    timestep = t_cast.to(dtype).to(x0.device)
    index = int(torch.argmin((timesteps - t_cast.cpu()).abs()))
    sigma = float(sigmas[index])
    latents = (1 - sigma) * x0 + sigma * noise                                              # add_noise, FM:164-170
    target = noise - x0                                                                     # training_target, FM:172-175
    ww = dict(w)
    ww["__linear__"] = stage2_linear_fn(lora, b2, masks) if b2 is not None else stage1_linear_fn(lora, masks)   # b2 None: stage 1
    pred = o.dit_forward(ww, cfg, latents.to(dtype), timestep, context, fuse_vae_embedding_in_latents)  # LOSS:17
    loss = F.mse_loss(pred.float(), target.float()) * float(weights[index])                   # LOSS:19-20
    return (loss, pred) if return_pred else loss


def loss_and_grads(w: Weights, cfg: o.DiTConfig, lora: Weights, b2: Weights, masks, x0, noise, timestep_id, context,
                   fuse_vae_embedding_in_latents: bool = True, timestep_dtype=None):
    """(loss, prediction, {module: dloss/dB2}) by autograd — the quantities of one reference training step."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in b2.items()}
    loss, pred = sft_loss(w, cfg, lora, leaves, masks, x0, noise, timestep_id, context, fuse_vae_embedding_in_latents,
                          return_pred=True, timestep_dtype=timestep_dtype)
    loss.backward()
    return loss.detach(), pred.detach(), {k: v.grad.detach() for k, v in leaves.items()}


def loss_and_grads_stage1(w: Weights, cfg: o.DiTConfig, lora: Weights, masks, x0, noise, timestep_id, context,
                          fuse_vae_embedding_in_latents: bool = True, timestep_dtype=None):
    """Stage 1 (identity LoRA): (loss, prediction, {lora key: dloss/d(A or B)}) by autograd."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in lora.items()}
    loss, pred = sft_loss(w, cfg, leaves, None, masks, x0, noise, timestep_id, context, fuse_vae_embedding_in_latents,
                          return_pred=True, timestep_dtype=timestep_dtype)
    loss.backward()
    return loss.detach(), pred.detach(), {k: v.grad.detach() for k, v in leaves.items()}
