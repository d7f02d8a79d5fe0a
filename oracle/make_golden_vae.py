"""Generate tests/golden/vae38.npz by running the REAL reference VAE38 (read-only, /root/reference/animation) at reduced
widths (dec_dim 16, z_dim 8 — the module's own constructor arguments) on seeded weights and latents.  Build container only.

    python oracle/make_golden_vae.py            # rewrites tests/golden/vae38.npz
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
from transformers import AutoTokenizer  # noqa: F401

for _m in ["imageio", "imageio.v3", "peft", "accelerate", "modelscope", "ftfy", "xfuser", "xfuser.core",
           "xfuser.core.distributed", "xfuser.core.long_ctx_attention"]:
    sys.modules[_m] = MagicMock()

from diffsynth.models import wan_video_vae as rv  # noqa: E402

from oracle import vae38_oracle as o  # noqa: E402


def latents(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


@torch.no_grad()
def main():
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    model = rv.VideoVAE38_(dim=cfg.enc_dim, z_dim=cfg.z_dim, dec_dim=cfg.dec_dim).eval()
    ref_sd = model.state_dict()
    dec_keys = {k: tuple(v.shape) for k, v in ref_sd.items() if k.startswith("decoder.") or k.startswith("conv2.")}
    assert dec_keys == o.param_shapes(cfg), set(dec_keys) ^ set(o.param_shapes(cfg))
    assert o.count_cache_slots(cfg) == rv.count_conv3d(model.decoder)
    enc_keys = {k: tuple(v.shape) for k, v in ref_sd.items() if k.startswith("encoder.") or k.startswith("conv1.")}
    assert enc_keys == o.enc_param_shapes(cfg), set(enc_keys) ^ set(o.enc_param_shapes(cfg))
    assert o.count_enc_cache_slots(cfg) == rv.count_conv3d(model.encoder)
    sd = dict(w)
    sd.update(o.make_enc_weights(cfg, seed=0))
    assert set(sd) == set(ref_sd)                      # decoder + conv2 + encoder + conv1 = the whole module
    model.load_state_dict(sd, strict=True)
    vae = object.__new__(rv.WanVideoVAE38)          # the wrapper's methods without building the full-width default model
    torch.nn.Module.__init__(vae)
    mean, inv_std = o.latent_scale(cfg)
    vae.mean, vae.std, vae.scale = mean, 1.0 / inv_std, [mean, inv_std]
    vae.model, vae.upsampling_factor, vae.z_dim = model, 16, cfg.z_dim
    out = {}
    z = latents((1, cfg.z_dim, 3, 3, 4), 1)
    out["model_decode"] = model.decode(z, vae.scale).numpy()                      # [1, 3, 9, 48, 64], not clamped
    out["single_frame"] = vae.decode(latents((1, cfg.z_dim, 1, 2, 2), 2) * 3, device="cpu").numpy()   # clamped image
    zt = latents((1, cfg.z_dim, 2, 5, 5), 3)
    out["tiled"] = vae.decode(zt, device="cpu", tiled=True, tile_size=(3, 3), tile_stride=(2, 2)).float().numpy()
    out["tiled_ragged"] = vae.decode(latents((1, cfg.z_dim, 1, 4, 7), 4), device="cpu", tiled=True, tile_size=(3, 4),
                                     tile_stride=(2, 3)).float().numpy()
    # encoder (first-frame conditioning and a 9-frame clip: chunks of 1, 4, 4 frames)
    img = torch.tanh(latents((3, 1, 32, 48), 30))
    out["encode_image"] = vae.encode([img], device="cpu").numpy()                                   # [1, 8, 1, 2, 3]
    clip = torch.tanh(latents((3, 9, 32, 32), 31))
    out["encode_clip"] = vae.encode([clip], device="cpu").numpy()                                   # [1, 8, 3, 2, 2]
    out["encode_tiled"] = vae.encode([torch.tanh(latents((3, 1, 80, 96), 32))], device="cpu", tiled=True, tile_size=(3, 4),
                                     tile_stride=(2, 2)).numpy()                                    # [1, 8, 1, 5, 6]
    xa = latents((1, 16, 3, 4, 6), 33)
    out["avg_down_t2"] = rv.AvgDown3D(16, 32, factor_t=2, factor_s=2)(xa).numpy()
    out["avg_down_t1"] = rv.AvgDown3D(16, 16, factor_t=1, factor_s=2)(xa).numpy()
    out["patchify"] = rv.patchify(latents((1, 3, 2, 6, 10), 34), 2).numpy()
    # primitives
    x = latents((1, 32, 1, 3, 4), 5)
    up = rv.DupUp3D(32, 16, factor_t=2, factor_s=2)
    out["dup_t2_first"], out["dup_t2"] = up(x, True).numpy(), up(x, False).numpy()
    out["dup_t1"] = rv.DupUp3D(32, 32, factor_t=1, factor_s=2)(x, True).numpy()
    out["unpatchify"] = rv.unpatchify(latents((1, 12, 2, 3, 5), 6), 2).numpy()
    blk = model.decoder.middle[1]
    out["attention"] = blk(latents((1, cfg.dims[0], 2, 3, 4), 7)).numpy()
    res = model.decoder.upsamples[0].upsamples[3]                                # Resample38 'upsample3d'
    cache, chunks = [None], []
    for i in range(3):
        chunks.append(res(latents((1, cfg.dims[1], 1, 2, 3), 10 + i), cache, [0]).numpy())
    out["resample3d_c0"], out["resample3d_c1"], out["resample3d_c2"] = chunks
    rb = model.decoder.upsamples[2].upsamples[0]                                 # ResidualBlock with a 1x1x1 shortcut
    cache, chunks = [None, None], []
    for i in range(3):
        chunks.append(rb(latents((1, cfg.dims[2], 1 if i < 2 else 2, 3, 3), 20 + i), cache, [0]).numpy())
    out["resblock_c0"], out["resblock_c1"], out["resblock_c2"] = chunks
    path = os.path.join(REPO, "tests", "golden", "vae38.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
