"""Import the UNMODIFIED reference (``diffsynth`` as installed by ``baseline/install_ref.py`` into ``baseline/_ref``)
and build its own objects — ``WanModel``, ``WanVideoPipeline`` — for the reference arm of ``bench.py`` and for the
parity tests.  Test / measurement infrastructure only: nothing under ``fairygen_b200/`` imports this.

Import recipe (verified): ``transformers`` first, then stub the packages the reference imports at module
scope but never touches on the TI2V-5B DiT path and that the image does not have (imageio, peft, accelerate, modelscope,
ftfy, xfuser).  No reference code is edited; the only switch flipped is the reference's own module-level availability
flag ``wan_video_dit.FLASH_ATTN_2_AVAILABLE`` (DIT:14-18) when flash-attn cannot run on the device at hand, which is the
fallback the reference takes by itself on a machine without flash-attn (DIT:54-59, in-tree SDPA).
"""
from __future__ import annotations

import os
import sys
import types
from typing import Dict, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_STUBS = ["imageio", "imageio.v3", "peft", "accelerate", "modelscope", "ftfy", "xfuser", "xfuser.core",
          "xfuser.core.distributed", "xfuser.core.long_ctx_attention"]
_loaded: Optional[types.SimpleNamespace] = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "diffsynth", "pipelines", "wan_video.py"))


def load() -> types.SimpleNamespace:
    """Namespace of reference modules: wv (pipelines.wan_video), wd (models.wan_video_dit), fm (diffusion.flow_match),
    lora (utils.lora.general)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("baseline/_ref is empty: run `python baseline/install_ref.py` in the build container")
    from unittest.mock import MagicMock

    import transformers  # noqa: F401  must precede the stubs (its find_spec("accelerate") chokes on a mock)
    from transformers import AutoTokenizer, Wav2Vec2Processor  # noqa: F401

    for m in _STUBS:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import diffsynth.pipelines.wan_video as wv
    from diffsynth.diffusion import flow_match as fm
    from diffsynth.models import wan_video_dit as wd
    from diffsynth.utils.lora import general as lora

    _loaded = types.SimpleNamespace(wv=wv, wd=wd, fm=fm, lora=lora)
    return _loaded


def select_attention_backend(device) -> str:
    """Which branch of the reference's flash_attention() (DIT:27-60) will run on `device`, after checking that flash-attn 2
    really launches there (the wheel in this image has no sm_100 cubins on some builds).  Returns "flash_attn_2" or "sdpa"."""
    import torch

    ref = load()
    wd = ref.wd
    dev = torch.device(device)
    if dev.type != "cuda":
        wd.FLASH_ATTN_2_AVAILABLE = False
        wd.FLASH_ATTN_3_AVAILABLE = False
        return "sdpa"
    if wd.FLASH_ATTN_3_AVAILABLE:
        return "flash_attn_3"
    if wd.FLASH_ATTN_2_AVAILABLE:
        try:
            q = torch.randn(1, 256, 2, 128, device=dev, dtype=torch.bfloat16)
            out = wd.flash_attn.flash_attn_func(q, q, q)
            torch.cuda.synchronize(dev)
            ref_out = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), q.transpose(1, 2), q.transpose(1, 2)).transpose(1, 2)
            if not torch.isfinite(out.float()).all() or (out.float() - ref_out.float()).abs().max() > 5e-2:
                raise RuntimeError("flash-attn 2 result mismatch")
            return "flash_attn_2"
        except Exception:
            wd.FLASH_ATTN_2_AVAILABLE = False
    return "sdpa"


def build_wan_model(cfg, state_dict: Optional[Dict] = None, device="cpu", dtype=None, fill: Optional[str] = None):
    """The reference's ``WanModel`` (DIT:271-336) with TI2V-5B style flags and the dimensions of `cfg`
    (a fairygen_b200.WanDiTConfig or oracle DiTConfig: dim, in_dim, ffn_dim, out_dim, text_dim, freq_dim, eps, patch_size,
    num_heads, num_layers).  state_dict: tensors under the reference's own names (loaded with assign=True, so device /
    dtype come from them).  fill="tiled": parameters allocated uninitialised and filled from one small random tile —
    for TIMING runs of the 5 B-parameter model, where the default init alone takes a minute of single-threaded RNG."""
    import torch

    ref = load()
    kw = dict(dim=cfg.dim, in_dim=cfg.in_dim, ffn_dim=cfg.ffn_dim, out_dim=cfg.out_dim, text_dim=cfg.text_dim,
              freq_dim=cfg.freq_dim, eps=cfg.eps, patch_size=tuple(cfg.patch_size), num_heads=cfg.num_heads,
              num_layers=cfg.num_layers, has_image_input=False, seperated_timestep=True, require_clip_embedding=False,
              require_vae_embedding=False, fuse_vae_embedding_in_latents=True)
    if state_dict is not None or fill is not None:
        with torch.device("meta"):
            dit = ref.wd.WanModel(**kw)
        # `freqs` is a plain attribute (tuple of complex tables), built on the meta device above: rebuild it for real
        dit.freqs = ref.wd.precompute_freqs_cis_3d(cfg.dim // cfg.num_heads)
        if state_dict is not None:
            missing, unexpected = dit.load_state_dict(state_dict, strict=True, assign=True)
            assert not missing and not unexpected
        else:
            dit.to_empty(device=device)
            g = torch.Generator(device="cpu").manual_seed(0)
            tile = (torch.rand(1 << 20, generator=g) * 2 - 1).to(device=device, dtype=dtype or torch.float32)
            with torch.no_grad():
                for name, p in dit.named_parameters():
                    n = p.numel()
                    scale = 1.0 / max(1.0, float(p.shape[-1] if p.dim() > 1 else 1)) ** 0.5
                    flat = p.data.view(-1)
                    if "norm" in name and name.endswith("weight"):
                        flat.fill_(1.0)
                    else:
                        reps = -(-n // tile.numel())
                        flat.copy_((tile.repeat(reps)[:n] * scale).to(p.dtype))
    else:
        dit = ref.wd.WanModel(**kw)
    dit = dit.eval()
    if dtype is not None and state_dict is None:
        dit = dit.to(dtype)
    if state_dict is None and fill is None:
        dit = dit.to(device)
    for p in dit.parameters():
        p.requires_grad_(False)
    return dit


class _StubTokenizer:
    """Stands in for the reference's HuggingfaceTokenizer (vocabulary files are not in the tree): a prompt becomes
    `len(prompt.split())` live ids (at least 1) in a `text_len`-long row, mask = 1 on the live prefix."""

    def __init__(self, text_len: int):
        self.text_len = text_len

    def __call__(self, prompt, return_mask=True, add_special_tokens=True):
        import torch

        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        ids = torch.zeros(len(prompts), self.text_len, dtype=torch.long)
        mask = torch.zeros(len(prompts), self.text_len, dtype=torch.long)
        for i, p in enumerate(prompts):
            n = max(1, min(self.text_len, len(p.split())))
            ids[i, :n] = torch.tensor([(hash_str(w) % 997) + 1 for w in (p.split() or ["_"])][:n])
            mask[i, :n] = 1
        return (ids, mask) if return_mask else ids


def hash_str(s: str) -> int:
    import zlib

    return zlib.crc32(s.encode())


def make_stub_text_encoder(text_dim: int, device, dtype):
    """Embedding table standing in for umT5 (the encoder is a separate §8(f) row with its own tests)."""
    import torch

    class StubTextEncoder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(77)
            self.table = torch.nn.Parameter(torch.randn(1000, text_dim, generator=g), requires_grad=False)

        def forward(self, ids, mask=None):
            return self.table[ids]

    return StubTextEncoder().to(device=device, dtype=dtype)


def make_stub_vae(z_dim: int, device, dtype):
    """Object with the attributes the pipeline touches (PIPE:347-360, 490-497, 322-323): `.model.z_dim`,
    `.upsampling_factor`, `.encode(list of (C,T,H,W))`, `.decode(latents)`; average-pool 'encode', nearest 'decode'."""
    import torch

    class StubVAE(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = types.SimpleNamespace(z_dim=z_dim)
            self.upsampling_factor = 16
            g = torch.Generator().manual_seed(78)
            self.mix = torch.nn.Parameter(torch.randn(z_dim, 3, generator=g), requires_grad=False)

        def encode(self, videos, device=None, tiled=False, tile_size=None, tile_stride=None):
            outs = []
            for v in videos:   # (3, T, H, W)
                x = torch.nn.functional.avg_pool2d(v.float().transpose(0, 1), 16).transpose(0, 1)   # (3, T, H/16, W/16)
                outs.append(torch.einsum("zc,cthw->zthw", self.mix.float(), x)[:, :1])
            return torch.stack(outs).to(dtype=dtype)

        def decode(self, latents, device=None, tiled=False, tile_size=None, tile_stride=None):
            self.last_latents = latents.detach().clone()     # what the denoise loop handed to the decoder
            x = latents[:, :3].float()
            x = x.repeat_interleave(4, dim=2)[:, :, 3:]
            return torch.nn.functional.interpolate(x, scale_factor=(1, 16, 16), mode="nearest").clamp(-1, 1)

    return StubVAE().to(device=device, dtype=dtype)


def build_pipeline(dit, device, dtype, text_len: int = 32, text_dim: Optional[int] = None):
    """The reference's own ``WanVideoPipeline`` object (PIPE:30-83) around `dit`, with the models that are not on the DiT hot
    path replaced by small stand-ins (tokenizer, text encoder, VAE).  ``pipe(...)`` then executes the UNMODIFIED
    ``__call__`` (PIPE:172-329): units, scheduler, the denoising loop and its ``self.model_fn`` calls."""
    ref = load()
    pipe = ref.wv.WanVideoPipeline(device=device, torch_dtype=dtype)
    pipe.dit = dit
    if text_dim is None:
        text_dim = dit.text_embedding[0].in_features
    pipe.tokenizer = _StubTokenizer(text_len)
    pipe.text_encoder = make_stub_text_encoder(text_dim, device, dtype)
    pipe.vae = make_stub_vae(dit.in_dim, device, dtype)
    return pipe
