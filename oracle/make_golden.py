"""Generate tests/golden/*.npz by running the REAL reference (read-only, /root/reference/animation).

Runs only in the build container (the reference tree does not exist on the GPU box).  The vectors
pin ``oracle/wan_dit_oracle.py`` — and through it the CUDA path — to the reference's behaviour.
Weights and inputs are regenerated from seeds by ``wan_dit_oracle.make_weights/make_inputs`` so
only OUTPUTS are stored.  Import recipe for the reference: SURVEY.md §8(c).

    python oracle/make_golden.py            # rewrites tests/golden/
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/animation"
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import transformers  # noqa: F401  (must precede the mocks)
from transformers import AutoTokenizer, Wav2Vec2Processor  # noqa: F401

for _m in ["imageio", "imageio.v3", "peft", "accelerate", "modelscope", "ftfy", "xfuser", "xfuser.core",
           "xfuser.core.distributed", "xfuser.core.long_ctx_attention"]:
    sys.modules[_m] = MagicMock()

import diffsynth.pipelines.wan_video as wv  # noqa: E402
from diffsynth.diffusion.flow_match import FlowMatchScheduler  # noqa: E402
from diffsynth.models import wan_video_dit as wd  # noqa: E402
from diffsynth.utils.lora.general import GeneralLoRALoader  # noqa: E402
from diffsynth.utils.xfuser import xdit_context_parallel as usp  # noqa: E402

from oracle import wan_dit_oracle as o  # noqa: E402

wd.FLASH_ATTN_2_AVAILABLE = False  # CPU: take the in-tree SDPA branch (DIT:54-59)
OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.manual_seed(1234)


def build_ref_model(cfg, weights, dtype):
    dit = wd.WanModel(dim=cfg.dim, in_dim=cfg.in_dim, ffn_dim=cfg.ffn_dim, out_dim=cfg.out_dim, text_dim=cfg.text_dim,
                      freq_dim=cfg.freq_dim, eps=cfg.eps, patch_size=cfg.patch_size, num_heads=cfg.num_heads,
                      num_layers=cfg.num_layers, has_image_input=False, seperated_timestep=True,
                      require_clip_embedding=False, require_vae_embedding=False,
                      fuse_vae_embedding_in_latents=True).eval()
    dit.load_state_dict(weights, strict=True)
    return dit.to(dtype)


def np32(t):
    return t.detach().float().cpu().numpy()


@torch.no_grad()
def golden_tiny_forward():
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    shape = (1, cfg.in_dim, 3, 8, 8)  # grid (3,4,4): S = 48 tokens, 16 first-frame tokens
    lat, z0, cp, cn = o.make_inputs(cfg, shape, text_len=32, live_text=8)
    dit = build_ref_model(cfg, w, torch.float32)
    out = {}
    for ts_val in (900.0, 37.0):
        ts = torch.tensor([ts_val])
        out[f"fused_t{int(ts_val)}"] = np32(wv.model_fn_wan_video(dit=dit, latents=lat, timestep=ts, context=cp,
                                                                 fuse_vae_embedding_in_latents=True))
    out["plain_t900"] = np32(wv.model_fn_wan_video(dit=dit, latents=lat, timestep=torch.tensor([900.0]), context=cn,
                                                   fuse_vae_embedding_in_latents=False))
    # reference bf16 path (CPU kernels): the "reference diffsynth bf16 path" at tiny scale
    dit16 = build_ref_model(cfg, w, torch.bfloat16)
    out["fused_t900_bf16"] = np32(wv.model_fn_wan_video(dit=dit16, latents=lat.bfloat16(),
                                                        timestep=torch.tensor([900.0]).bfloat16(), context=cp.bfloat16(),
                                                        fuse_vae_embedding_in_latents=True))
    # a ragged grid: (2,3,5) -> S = 30 (not a multiple of anything convenient)
    shape2 = (1, cfg.in_dim, 2, 6, 10)
    lat2, _, cp2, _ = o.make_inputs(cfg, shape2, text_len=24, live_text=24)
    out["ragged_fused_t500"] = np32(wv.model_fn_wan_video(dit=dit, latents=lat2, timestep=torch.tensor([500.0]),
                                                          context=cp2, fuse_vae_embedding_in_latents=True))
    np.savez_compressed(os.path.join(OUT, "tiny_forward.npz"), **out)
    print("tiny_forward:", {k: v.shape for k, v in out.items()})


@torch.no_grad()
def golden_ops():
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    dit = build_ref_model(cfg, w, torch.float32)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 30, cfg.dim, generator=g)
    out = {"x": np32(x)}
    out["sinusoid"] = np32(wd.sinusoidal_embedding_1d(256, torch.tensor([0.0, 1.0, 37.0, 500.0, 996.0, 1000.0])))
    out["sinusoid_bf16"] = np32(wd.sinusoidal_embedding_1d(256, torch.tensor([0.0, 996.0, 1000.0]).bfloat16()))
    f, h, wd_ = 2, 3, 5
    freqs = torch.cat([dit.freqs[0][:f].view(f, 1, 1, -1).expand(f, h, wd_, -1),
                       dit.freqs[1][:h].view(1, h, 1, -1).expand(f, h, wd_, -1),
                       dit.freqs[2][:wd_].view(1, 1, wd_, -1).expand(f, h, wd_, -1)], dim=-1).reshape(f * h * wd_, 1, -1)
    out["freqs_real"] = freqs.real.numpy()
    out["freqs_imag"] = freqs.imag.numpy()
    out["rope"] = np32(wd.rope_apply(x, freqs, cfg.num_heads))
    blk = dit.blocks[1]
    out["rmsnorm"] = np32(blk.self_attn.norm_q(x))
    out["modulate"] = np32(wd.modulate(x, x.flip(1) * 0.1, x.flip(2) * 0.2))
    ctx = torch.randn(1, 24, cfg.dim, generator=g)
    t_mod_tok = torch.randn(1, 30, 6, cfg.dim, generator=g) * 0.2
    t_mod_one = torch.randn(1, 6, cfg.dim, generator=g) * 0.2
    out["ctx"] = np32(ctx)
    out["t_mod_tok"] = np32(t_mod_tok)
    out["t_mod_one"] = np32(t_mod_one)
    out["self_attn"] = np32(blk.self_attn(x, freqs))
    out["cross_attn"] = np32(blk.cross_attn(x, ctx))
    out["block_tok"] = np32(blk(x, ctx, t_mod_tok, freqs))
    out["block_one"] = np32(blk(x, ctx, t_mod_one, freqs))
    t_tok = torch.randn(1, 30, cfg.dim, generator=g) * 0.3
    t_one = torch.randn(1, cfg.dim, generator=g) * 0.3
    out["t_tok"] = np32(t_tok)
    out["t_one"] = np32(t_one)
    out["head_tok"] = np32(dit.head(x, t_tok))
    out["head_one"] = np32(dit.head(x, t_one))
    out["unpatchify"] = np32(dit.unpatchify(dit.head(x, t_tok), (f, h, wd_)))
    lat = torch.randn(1, cfg.in_dim, 2, 6, 10, generator=g)
    out["lat"] = np32(lat)
    out["patchify"] = np32(dit.patchify(lat))
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **out)
    print("ops:", sorted(out))


@torch.no_grad()
def golden_scheduler():
    sch = FlowMatchScheduler("Wan")
    out = {}
    for n, shift in ((50, 5.0), (8, 3.0)):
        sch.set_timesteps(n, denoising_strength=1.0, shift=shift)
        out[f"sigmas_{n}"] = sch.sigmas.numpy()
        out[f"timesteps_{n}"] = sch.timesteps.numpy()
        out[f"timesteps_bf16_{n}"] = np32(sch.timesteps.to(torch.bfloat16))  # PIPE:293
    sch.set_timesteps(50, denoising_strength=1.0, shift=5.0)
    g = torch.Generator().manual_seed(5)
    sample = torch.randn(1, 48, 3, 8, 8, generator=g).bfloat16()
    npos = torch.randn(1, 48, 3, 8, 8, generator=g).bfloat16()
    nneg = torch.randn(1, 48, 3, 8, 8, generator=g).bfloat16()
    z0 = torch.randn(1, 48, 1, 8, 8, generator=g).bfloat16()
    out["sample"], out["npos"], out["nneg"], out["z0"] = np32(sample), np32(npos), np32(nneg), np32(z0)
    for i in (0, 17, 49):
        npred = nneg + 5.0 * (npos - nneg)  # PIPE:302, bf16 tensor ops
        nxt = sch.step(npred, sch.timesteps[i], sample)  # PIPE:307
        assert nxt.dtype == torch.bfloat16
        nxt[:, :, 0:1] = z0  # PIPE:308-309
        out[f"step_{i}"] = np32(nxt)
    np.savez_compressed(os.path.join(OUT, "scheduler.npz"), **out)
    print("scheduler:", sorted(out))


@torch.no_grad()
def golden_lora():
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    lora = o.make_lora(cfg, rank=8, seed=2)
    out = {}
    for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        dit = build_ref_model(cfg, w, dtype)
        GeneralLoRALoader(device="cpu", torch_dtype=dtype).fuse_lora_to_base_model(dit, lora, alpha=1.0)
        sd = dit.state_dict()
        for name in ("blocks.0.self_attn.q.weight", "blocks.1.cross_attn.o.weight", "blocks.1.ffn.2.weight",
                     "blocks.0.ffn.0.bias", "head.head.weight"):
            full = np32(sd[name])
            out[f"{tag}:{name}"] = full[:32]  # first rows + a checksum of the whole tensor keep the fixture small
            out[f"{tag}:{name}:sum"] = np.array(full.astype(np.float64).sum())
    np.savez_compressed(os.path.join(OUT, "lora.npz"), **out)
    print("lora:", sorted(out))


@torch.no_grad()
def golden_usp():
    """Sequence-parallel glue that IS in the tree (USP:30-55 rope with ones-padding; PIPE:1312-1315 chunk+pad)."""
    cfg = o.TINY
    g = torch.Generator().manual_seed(21)
    world, s_total = 4, 30  # chunk -> 8,8,8,6 ; pad 2
    x = torch.randn(1, s_total, cfg.dim, generator=g)
    tables = o.rope_tables_3d(cfg.head_dim)
    freqs = o.rope_freqs(tables, 2, 3, 5)
    chunks = torch.chunk(x, world, dim=1)
    chunks = [torch.nn.functional.pad(c, (0, 0, 0, chunks[0].shape[1] - c.shape[1]), value=0) for c in chunks]
    out = {"x": np32(x)}
    for rank in range(world):
        usp.get_sequence_parallel_world_size = lambda: world
        usp.get_sequence_parallel_rank = lambda r=rank: r
        out[f"rope_rank{rank}"] = np32(usp.rope_apply(chunks[rank], freqs, cfg.num_heads))
        out[f"chunk_rank{rank}"] = np32(chunks[rank])
    np.savez_compressed(os.path.join(OUT, "usp.npz"), **out)
    print("usp:", sorted(out))


@torch.no_grad()
def golden_denoise():
    """4-step CFG denoise of the tiny model in fp32 through the reference's model_fn + scheduler,
    with the loop body of PIPE:285-309 (the pipeline object itself needs a VAE/text encoder)."""
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    dit = build_ref_model(cfg, w, torch.float32)
    shape = (1, cfg.in_dim, 3, 8, 8)
    lat, z0, cp, cn = o.make_inputs(cfg, shape, text_len=32, live_text=8)
    lat[:, :, 0:1] = z0  # PIPE:496
    sch = FlowMatchScheduler("Wan")
    sch.set_timesteps(4, denoising_strength=1.0, shift=5.0)
    for i, ts in enumerate(sch.timesteps):
        t_in = ts.unsqueeze(0).to(dtype=torch.float32)
        npos = wv.model_fn_wan_video(dit=dit, latents=lat, timestep=t_in, context=cp, fuse_vae_embedding_in_latents=True)
        nneg = wv.model_fn_wan_video(dit=dit, latents=lat, timestep=t_in, context=cn, fuse_vae_embedding_in_latents=True)
        npred = nneg + 5.0 * (npos - nneg)
        lat = sch.step(npred, sch.timesteps[i], lat)
        lat[:, :, 0:1] = z0
    np.savez_compressed(os.path.join(OUT, "denoise.npz"), final=np32(lat))
    print("denoise: final mean|x| =", lat.abs().mean().item())


if __name__ == "__main__":
    golden_tiny_forward()
    golden_ops()
    golden_scheduler()
    golden_lora()
    golden_usp()
    golden_denoise()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"golden fixtures: {total/1024:.0f} KiB in {OUT}")
