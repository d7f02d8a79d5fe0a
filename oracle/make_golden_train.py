"""Generate tests/golden/train.npz by running the REAL reference DiT (read-only, /root/reference/animation) through
one stage-2 fine-tune step.  Build container only.

``peft`` is absent (and unpinned by the reference), so the adapted Linears of the real ``WanModel`` are replaced by
``Stage2Linear`` below, whose forward is TMOD:317-352 (animation/diffsynth/diffusion/training_module.py) line by
line with the Bernoulli mask injected; everything else — model_fn_wan_video, DiTBlock, the scheduler's add_noise /
training_target / training_weight, the loss expression of diffusion/loss.py:19-20 and autograd — is the reference's
own code.  The stored loss / prediction / B2 gradients pin ``oracle/wan_train_oracle.py``.

    python oracle/make_golden_train.py
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/animation"
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import transformers  # noqa: F401,E402
from transformers import AutoTokenizer, Wav2Vec2Processor  # noqa: F401,E402

for _m in ["imageio", "imageio.v3", "peft", "accelerate", "modelscope", "ftfy", "xfuser", "xfuser.core",
           "xfuser.core.distributed", "xfuser.core.long_ctx_attention"]:
    sys.modules[_m] = MagicMock()

import diffsynth.pipelines.wan_video as wv  # noqa: E402
from diffsynth.diffusion.flow_match import FlowMatchScheduler  # noqa: E402
from diffsynth.models import wan_video_dit as wd  # noqa: E402

from oracle import wan_dit_oracle as o  # noqa: E402
from oracle import wan_train_oracle as t  # noqa: E402

wd.FLASH_ATTN_2_AVAILABLE = False
OUT = os.path.join(REPO, "tests", "golden")


class Stage2Linear(torch.nn.Module):
    """TMOD:317-352 (`new_forward` of stage-2) on a plain nn.Linear base layer."""

    def __init__(self, base, a1, b1, b2, mask):
        super().__init__()
        self.base_layer = base
        self.a1, self.b1, self.mask = a1, b1, mask
        self.lora_B2 = torch.nn.Parameter(b2.clone())
        self.scaling = 1.0

    def forward(self, x):
        result = self.base_layer(x)
        result = result + F.linear(F.linear(x, self.a1), self.b1) * self.scaling
        dropout_prob = 0.5
        mask = self.mask.to(self.lora_B2.dtype)
        scale_factor = 1.0 / (1 - dropout_prob)
        b2_dropped = self.lora_B2 * mask * scale_factor
        intermediate = F.linear(x, self.a1)
        update = F.linear(intermediate, b2_dropped, None)
        return result + update * self.scaling


class Stage1Linear(torch.nn.Module):
    """TMOD:200-264 (`new_forward` of stage-1) on a plain nn.Linear base layer; A and B trainable."""

    def __init__(self, base, a, b, mask):
        super().__init__()
        self.base_layer = base
        self.lora_A = torch.nn.Parameter(a.clone())
        self.lora_B = torch.nn.Parameter(b.clone())
        self.mask = mask
        self.scaling = 1.0

    def forward(self, x):
        result = self.base_layer(x)
        dropout_prob = 0.8
        mask = self.mask.to(self.lora_B.dtype)
        scale_factor = 1.0 / (1 - dropout_prob)
        b_dropped = self.lora_B * mask * scale_factor
        intermediate = F.linear(x, self.lora_A)
        update = F.linear(intermediate, b_dropped, None)
        return result + update * self.scaling


def build_ref(cfg, w):
    dit = wd.WanModel(dim=cfg.dim, in_dim=cfg.in_dim, ffn_dim=cfg.ffn_dim, out_dim=cfg.out_dim, text_dim=cfg.text_dim,
                      freq_dim=cfg.freq_dim, eps=cfg.eps, patch_size=cfg.patch_size, num_heads=cfg.num_heads,
                      num_layers=cfg.num_layers, has_image_input=False, seperated_timestep=True,
                      require_clip_embedding=False, require_vae_embedding=False, fuse_vae_embedding_in_latents=True).eval()
    dit.load_state_dict(w, strict=True)
    for p in dit.parameters():
        p.requires_grad_(False)
    return dit


def wrap(dit, cfg, factory):
    wrapped = {}
    for name in t.lora_targets(cfg):
        parent_name, child = name.rsplit(".", 1)
        parent = dit.get_submodule(parent_name)
        base = getattr(parent, child) if not child.isdigit() else parent[int(child)]
        mod = factory(name, base)
        if child.isdigit():
            parent[int(child)] = mod
        else:
            setattr(parent, child, mod)
        wrapped[name] = mod
    return wrapped


STAGE1_STORED = ("blocks.0.self_attn.q", "blocks.0.cross_attn.k", "blocks.0.ffn.0", "blocks.1.self_attn.o", "blocks.1.cross_attn.v",
                 "blocks.1.ffn.2")


def golden_stage1(out):
    """Stage-1 (identity) LoRA step: A and B trainable, weight dropout 0.8 (TMOD:200-264)."""
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    lora = o.make_lora(cfg, rank=32, seed=2)
    masks = t.make_masks_stage1(cfg, rank=32)
    dit = build_ref(cfg, w)
    wrapped = wrap(dit, cfg, lambda name, base: Stage1Linear(base, lora[f"{name}.lora_A.default.weight"],
                                                            lora[f"{name}.lora_B.default.weight"], masks[name]))
    sched = FlowMatchScheduler("Wan")
    sched.set_timesteps(1000, training=True)
    shape, text_len, timestep_id = (1, cfg.in_dim, 3, 8, 8), 32, 500
    x0, _, ctx, _ = o.make_inputs(cfg, shape, text_len=text_len, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    timestep = sched.timesteps[torch.tensor([timestep_id])]
    latents = sched.add_noise(x0, noise, timestep)
    target = sched.training_target(x0, noise, timestep)
    pred = wv.model_fn_wan_video(dit=dit, latents=latents, timestep=timestep, context=ctx, fuse_vae_embedding_in_latents=True)
    loss = F.mse_loss(pred.float(), target.float()) * sched.training_weight(timestep)
    loss.backward()
    out["s1_loss"] = np.array(float(loss.detach()), dtype=np.float64)
    out["s1_pred"] = pred.detach().numpy()
    for name in STAGE1_STORED:
        out[f"s1_gradA.{name}"] = wrapped[name].lora_A.grad.detach().numpy().astype(np.float32)
        out[f"s1_gradB.{name}"] = wrapped[name].lora_B.grad.detach().numpy().astype(np.float32)
    print("stage1 loss", float(loss.detach()))


def main():
    cfg = o.TINY
    rank = 32
    w = o.make_weights(cfg, seed=0)
    lora = o.make_lora(cfg, rank=rank, seed=2)
    b2 = t.make_b2(cfg, rank=rank)
    masks = t.make_masks(cfg, rank=rank)
    dit = wd.WanModel(dim=cfg.dim, in_dim=cfg.in_dim, ffn_dim=cfg.ffn_dim, out_dim=cfg.out_dim, text_dim=cfg.text_dim,
                      freq_dim=cfg.freq_dim, eps=cfg.eps, patch_size=cfg.patch_size, num_heads=cfg.num_heads,
                      num_layers=cfg.num_layers, has_image_input=False, seperated_timestep=True,
                      require_clip_embedding=False, require_vae_embedding=False, fuse_vae_embedding_in_latents=True).eval()
    dit.load_state_dict(w, strict=True)
    for p in dit.parameters():
        p.requires_grad_(False)
    wrapped = {}
    for name in t.lora_targets(cfg):
        parent_name, child = name.rsplit(".", 1)
        parent = dit.get_submodule(parent_name)
        base = getattr(parent, child) if not child.isdigit() else parent[int(child)]
        mod = Stage2Linear(base, lora[f"{name}.lora_A.default.weight"], lora[f"{name}.lora_B.default.weight"], b2[name], masks[name])
        if child.isdigit():
            parent[int(child)] = mod
        else:
            setattr(parent, child, mod)
        wrapped[name] = mod

    sched = FlowMatchScheduler("Wan")
    sched.set_timesteps(1000, training=True)
    out = {}
    for tag, shape, text_len, timestep_id in (("a", (1, cfg.in_dim, 3, 8, 8), 32, 500), ("b", (1, cfg.in_dim, 2, 6, 10), 24, 37)):
        x0, _, ctx, _ = o.make_inputs(cfg, shape, text_len=text_len, live_text=8)
        noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
        for m in wrapped.values():
            m.lora_B2.grad = None
        timestep = sched.timesteps[torch.tensor([timestep_id])]
        latents = sched.add_noise(x0, noise, timestep)                      # LOSS:13
        target = sched.training_target(x0, noise, timestep)                 # LOSS:14
        pred = wv.model_fn_wan_video(dit=dit, latents=latents, timestep=timestep, context=ctx, fuse_vae_embedding_in_latents=True)
        loss = F.mse_loss(pred.float(), target.float()) * sched.training_weight(timestep)   # LOSS:19-20
        loss.backward()
        out[f"{tag}_loss"] = np.array(float(loss), dtype=np.float64)
        out[f"{tag}_pred"] = pred.detach().numpy()
        out[f"{tag}_weight"] = np.array(float(sched.training_weight(timestep)))
        out[f"{tag}_timestep"] = np.array(float(timestep))
        for name, m in wrapped.items():
            out[f"{tag}_grad.{name}"] = m.lora_B2.grad.detach().numpy().astype(np.float32)
        print(tag, "loss", float(loss), "timestep", float(timestep))
    sig, ts, wts = t.training_schedule()
    assert torch.equal(sig, sched.sigmas) and torch.equal(ts, sched.timesteps) and torch.allclose(wts, sched.linear_timesteps_weights)
    golden_stage1(out)
    np.savez_compressed(os.path.join(OUT, "train.npz"), **out)
    print("train.npz:", len(out), "arrays", os.path.getsize(os.path.join(OUT, "train.npz")), "bytes")


if __name__ == "__main__":
    main()
