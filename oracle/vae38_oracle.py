"""CPU/PyTorch ORACLE for the Wan2.2 VAE38 decoder and encoder.  TEST INFRASTRUCTURE ONLY.

A functional restatement (plain torch ops over a flat ``{name: tensor}`` weight dict, explicit feature-cache list) of what
the reference's ``WanVideoVAE38.decode`` computes: latent de-normalisation, ``conv2``, the chunk-by-chunk causal decoder
with its two-frame feature cache, ``unpatchify``, the tiled variant with linear-ramp blending, and the final clamp; and of
``WanVideoVAE38.encode`` (patchify, chunked causal encoder with AvgDown3D shortcuts, ``conv1``, latent normalisation, tiled
variant) — the encoder has no CUDA path yet, its oracle is the groundwork for it.  Not part of the product: only ``tests/``
and bench CPU-baseline legs may import it.

Parity status: PINNED.  ``oracle/make_golden_vae.py`` runs the real ``WanVideoVAE38`` (imported from
``/root/reference/animation`` in the build container) at reduced widths on seeded weights / latents and stores its outputs
in ``tests/golden/vae38.npz``; ``tests/test_vae_oracle.py`` checks this file against them.

Reference file restated (relative to /root/reference/animation/diffsynth):  VAE = models/wan_video_vae.py
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]
CACHE_T = 2                                               # VAE:8

# WanVideoVAE38.__init__, VAE:1359-1376
MEAN38 = [-0.2289, -0.0052, -0.1323, -0.2339, -0.2799, 0.0174, 0.1838, 0.1557, -0.1382, 0.0542, 0.2813, 0.0891, 0.1570, -0.0098, 0.0375,
          -0.1825, -0.2246, -0.1207, -0.0698, 0.5109, 0.2665, -0.2108, -0.2158, 0.2502, -0.2055, -0.0322, 0.1109, 0.1567, -0.0729, 0.0899,
          -0.2799, -0.1230, -0.0313, -0.1649, 0.0117, 0.0723, -0.2839, -0.2083, -0.0520, 0.3748, 0.0152, 0.1957, 0.1433, -0.2944, 0.3573,
          -0.0548, -0.1681, -0.0667]
STD38 = [0.4765, 1.0364, 0.4514, 1.1677, 0.5313, 0.4990, 0.4818, 0.5013, 0.8158, 1.0344, 0.5894, 1.0901, 0.6885, 0.6165, 0.8454, 0.4978,
         0.5759, 0.3523, 0.7135, 0.6804, 0.5833, 1.4146, 0.8986, 0.5659, 0.7069, 0.5338, 0.4889, 0.4917, 0.4069, 0.4999, 0.6866, 0.4093,
         0.5709, 0.6065, 0.6415, 0.4944, 0.5726, 1.2042, 0.5458, 1.6887, 0.3971, 1.0600, 0.3943, 0.5537, 0.5444, 0.4089, 0.7468, 0.7744]


@dataclass(frozen=True)
class VAE38Config:                                       # VideoVAE38_.__init__ defaults, VAE:1271-1279
    z_dim: int = 48
    dec_dim: int = 256
    dim_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temperal_upsample: Tuple[bool, ...] = (True, True, False)   # temperal_downsample[::-1], VAE:1288 (sic)
    out_channels: int = 12                               # 3 x 2 x 2, un-patchified to RGB at twice the size (VAE:887, 1350)
    enc_dim: int = 160                                   # `dim` of VideoVAE38_ / WanVideoVAE38 (VAE:1272, 1356)
    temperal_downsample: Tuple[bool, ...] = (False, True, True)   # VAE:1278

    @property
    def dims(self) -> List[int]:                         # VAE:859
        return [self.dec_dim * u for u in [self.dim_mult[-1]] + list(self.dim_mult[::-1])]

    @property
    def enc_dims(self) -> List[int]:                     # VAE:638
        return [self.enc_dim * u for u in [1] + list(self.dim_mult)]

    @property
    def upsampling_factor(self) -> int:                  # 2^(stages with up_flag) x patch 2  (= 16, VAE:1380)
        return 2 ** (len(self.dim_mult) - 1) * 2


VAE38 = VAE38Config()
TINY = VAE38Config(z_dim=8, dec_dim=16, enc_dim=16)


def stage_plan(cfg: VAE38Config):
    """(in_dim, out_dim, n residual blocks, up_flag, temporal up) of the Up_ResidualBlocks, VAE:869-879."""
    dims = cfg.dims
    plan = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        t_up = cfg.temperal_upsample[i] if i < len(cfg.temperal_upsample) else False
        plan.append((a, b, cfg.num_res_blocks + 1, i != len(cfg.dim_mult) - 1, t_up))
    return plan


def _res_shapes(p: str, cin: int, cout: int) -> Dict[str, tuple]:
    s = {p + "residual.0.gamma": (cin, 1, 1, 1), p + "residual.2.weight": (cout, cin, 3, 3, 3), p + "residual.2.bias": (cout,),
         p + "residual.3.gamma": (cout, 1, 1, 1), p + "residual.6.weight": (cout, cout, 3, 3, 3), p + "residual.6.bias": (cout,)}
    if cin != cout:
        s[p + "shortcut.weight"] = (cout, cin, 1, 1, 1)
        s[p + "shortcut.bias"] = (cout,)
    return s


def param_shapes(cfg: VAE38Config) -> Dict[str, tuple]:
    """State-dict keys / shapes of ``VideoVAE38_`` that decode touches: ``conv2`` and ``decoder`` (VAE:1294-1296, 842-887)."""
    d0 = cfg.dims[0]
    s = {"conv2.weight": (cfg.z_dim, cfg.z_dim, 1, 1, 1), "conv2.bias": (cfg.z_dim,),
         "decoder.conv1.weight": (d0, cfg.z_dim, 3, 3, 3), "decoder.conv1.bias": (d0,)}
    s.update(_res_shapes("decoder.middle.0.", d0, d0))
    s.update({"decoder.middle.1.norm.gamma": (d0, 1, 1), "decoder.middle.1.to_qkv.weight": (3 * d0, d0, 1, 1),
              "decoder.middle.1.to_qkv.bias": (3 * d0,), "decoder.middle.1.proj.weight": (d0, d0, 1, 1), "decoder.middle.1.proj.bias": (d0,)})
    s.update(_res_shapes("decoder.middle.2.", d0, d0))
    for i, (cin, cout, n, up, t_up) in enumerate(stage_plan(cfg)):
        p = f"decoder.upsamples.{i}.upsamples."
        c = cin
        for j in range(n):
            s.update(_res_shapes(f"{p}{j}.", c, cout))
            c = cout
        if up:
            s[f"{p}{n}.resample.1.weight"] = (cout, cout, 3, 3)
            s[f"{p}{n}.resample.1.bias"] = (cout,)
            if t_up:
                s[f"{p}{n}.time_conv.weight"] = (2 * cout, cout, 3, 1, 1)
                s[f"{p}{n}.time_conv.bias"] = (2 * cout,)
    c_last = cfg.dims[-1]
    s.update({"decoder.head.0.gamma": (c_last, 1, 1, 1), "decoder.head.2.weight": (cfg.out_channels, c_last, 3, 3, 3),
              "decoder.head.2.bias": (cfg.out_channels,)})
    return s


def make_weights(cfg: VAE38Config, seed: int = 0) -> Weights:
    """Seeded fp32 weights, tensor by tensor; variance-preserving scales so activations stay O(1) through ~40 layers."""
    out = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        if name.endswith("gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for v in shape[1:]:
                fan_in *= v
            t = torch.randn(shape, generator=g) * (fan_in ** -0.5)
        out[name] = t
    return out


def latent_scale(cfg: VAE38Config):
    """(mean, 1/std) of the latent channels: the real statistics for z_dim 48, their first z_dim entries otherwise."""
    return torch.tensor(MEAN38[:cfg.z_dim]), 1.0 / torch.tensor(STD38[:cfg.z_dim])


# --------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------
def causal_conv3d(x, weight, bias, cache_x=None):
    """CausalConv3d.forward, VAE:44-52: zero padding in h, w; 2*pad_t frames in front, the cached frames taking their place."""
    kt, kh, kw = weight.shape[2:]
    pad = [kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0]
    if cache_x is not None and pad[4] > 0:
        x = torch.cat([cache_x.to(x.device), x], dim=2)
        pad[4] -= cache_x.shape[2]
    return F.conv3d(F.pad(x, pad), weight, bias)


def rms_norm(x, gamma):
    """RMS_norm.forward (channel_first, no bias), VAE:67-70."""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


def _cache_tail(x, prev):
    """The last CACHE_T frames of x, topped up with the last cached frame when x is a single frame (VAE:288-295)."""
    c = x[:, :, -CACHE_T:].clone()
    if c.shape[2] < 2 and prev is not None:
        c = torch.cat([prev[:, :, -1:].to(c.device), c], dim=2)
    return c


def residual_block(w: Weights, p: str, x, cache: List, idx: List[int]):
    """ResidualBlock.forward with the feature cache, VAE:283-301."""
    h = causal_conv3d(x, w[p + "shortcut.weight"], w[p + "shortcut.bias"]) if (p + "shortcut.weight") in w else x
    for norm, conv in (("residual.0.", "residual.2."), ("residual.3.", "residual.6.")):
        x = F.silu(rms_norm(x, w[p + norm + "gamma"]))
        i = idx[0]
        tail = _cache_tail(x, cache[i])
        x = causal_conv3d(x, w[p + conv + "weight"], w[p + conv + "bias"], cache[i])
        cache[i] = tail
        idx[0] += 1
    return x + h


def attention_block(w: Weights, p: str, x):
    """AttentionBlock.forward, VAE:321-342: one head of width C over the h*w positions of each frame."""
    b, c, t, hh, ww = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)
    y = rms_norm(y, w[p + "norm.gamma"])
    qkv = F.conv2d(y, w[p + "to_qkv.weight"], w[p + "to_qkv.bias"]).reshape(b * t, 1, 3 * c, -1).permute(0, 1, 3, 2)
    q, k, v = qkv.chunk(3, dim=-1)
    y = F.scaled_dot_product_attention(q, k, v).squeeze(1).permute(0, 2, 1).reshape(b * t, c, hh, ww)
    y = F.conv2d(y, w[p + "proj.weight"], w[p + "proj.bias"])
    return y.reshape(b, t, c, hh, ww).permute(0, 2, 1, 3, 4) + x


def resample_up(w: Weights, p: str, x, temporal: bool, cache: List, idx: List[int]):
    """Resample38.forward for 'upsample2d' / 'upsample3d', VAE:120-160: the first chunk skips the temporal doubling
    ('Rep'), later chunks run time_conv on (cached frame | zeros, x) and interleave its two halves as frames; then
    nearest-exact x2 in h, w and a 3x3 Conv2d per frame."""
    b, c, t, hh, ww = x.shape
    if temporal:
        i = idx[0]
        if cache[i] is None:
            cache[i] = "Rep"
        else:
            tail = x[:, :, -CACHE_T:].clone()
            rep = isinstance(cache[i], str)
            if tail.shape[2] < 2:
                tail = torch.cat([torch.zeros_like(tail) if rep else cache[i][:, :, -1:], tail], dim=2)
            x = causal_conv3d(x, w[p + "time_conv.weight"], w[p + "time_conv.bias"], None if rep else cache[i])
            cache[i] = tail
            x = x.reshape(b, 2, c, t, hh, ww)
            x = torch.stack((x[:, 0], x[:, 1]), 3).reshape(b, c, t * 2, hh, ww)
        idx[0] += 1
    t = x.shape[2]
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)
    y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(y)
    y = F.conv2d(y, w[p + "resample.1.weight"], w[p + "resample.1.bias"], padding=1)
    return y.reshape(b, t, y.shape[1], 2 * hh, 2 * ww).permute(0, 2, 1, 3, 4)


def dup_up3d(x, out_channels: int, factor_t: int, factor_s: int, first_chunk: bool):
    """DupUp3D.forward, VAE:417-439: channel groups fan out to (t, h, w) sub-positions; parameter free."""
    repeats = out_channels * factor_t * factor_s * factor_s // x.shape[1]
    x = x.repeat_interleave(repeats, dim=1)
    b, _, t, hh, ww = x.shape
    x = x.view(b, out_channels, factor_t, factor_s, factor_s, t, hh, ww).permute(0, 1, 5, 2, 6, 3, 7, 4).contiguous()
    x = x.view(b, out_channels, t * factor_t, hh * factor_s, ww * factor_s)
    return x[:, :, factor_t - 1:] if first_chunk else x


def decoder_chunk(w: Weights, cfg: VAE38Config, x, cache: List, first_chunk: bool, trace=None):
    """Decoder3d_38.forward on one latent frame, VAE:889-940.  trace(name, tensor): intermediate results for the tests."""
    tr = trace if trace is not None else (lambda name, t: None)
    idx = [0]
    tail = _cache_tail(x, cache[0])
    x = causal_conv3d(x, w["decoder.conv1.weight"], w["decoder.conv1.bias"], cache[0])
    cache[0] = tail
    idx[0] = 1
    tr("conv1", x)
    x = residual_block(w, "decoder.middle.0.", x, cache, idx)
    tr("mid0", x)
    x = attention_block(w, "decoder.middle.1.", x)
    tr("attn", x)
    x = residual_block(w, "decoder.middle.2.", x, cache, idx)
    tr("mid2", x)
    for i, (cin, cout, n, up, t_up) in enumerate(stage_plan(cfg)):                    # Up_ResidualBlock.forward, VAE:506-514
        p = f"decoder.upsamples.{i}.upsamples."
        main = x
        for j in range(n):
            main = residual_block(w, f"{p}{j}.", main, cache, idx)
            tr(f"s{i}.b{j}", main)
        if up:
            main = resample_up(w, f"{p}{n}.", main, t_up, cache, idx)
            tr(f"s{i}.resample", main)
            x = main + dup_up3d(x, cout, 2 if t_up else 1, 2, first_chunk)
            tr(f"s{i}.out", x)
        else:
            x = main
    x = F.silu(rms_norm(x, w["decoder.head.0.gamma"]))
    i = idx[0]
    tail = _cache_tail(x, cache[i])
    x = causal_conv3d(x, w["decoder.head.2.weight"], w["decoder.head.2.bias"], cache[i])
    cache[i] = tail
    tr("head", x)
    return x


def count_cache_slots(cfg: VAE38Config) -> int:
    """count_conv3d(decoder), VAE:943-948: every CausalConv3d of the decoder, shortcuts and time_convs included."""
    return sum(1 for k, s in param_shapes(cfg).items() if k.startswith("decoder.") and k.endswith("weight") and len(s) == 5)


def unpatchify(x, patch: int = 2):
    """'b (c r q) f h w -> b c f (h q) (w r)', VAE:214-224."""
    b, crq, f, hh, ww = x.shape
    c = crq // (patch * patch)
    x = x.view(b, c, patch, patch, f, hh, ww)          # (c, r, q)
    return x.permute(0, 1, 4, 5, 3, 6, 2).reshape(b, c, f, hh * patch, ww * patch)


def model_decode(w: Weights, cfg: VAE38Config, z: torch.Tensor, trace=None) -> torch.Tensor:
    """VideoVAE38_.decode, VAE:1326-1351: z [1, z_dim, T, h, w] -> [1, 3, 4T-3, 16h, 16w] (not clamped)."""
    mean, inv_std = latent_scale(cfg)
    mean, inv_std = mean.to(z), inv_std.to(z)
    z = z / inv_std.view(1, -1, 1, 1, 1) + mean.view(1, -1, 1, 1, 1)
    x = causal_conv3d(z, w["conv2.weight"], w["conv2.bias"])
    cache: List = [None] * count_cache_slots(cfg)
    outs = [decoder_chunk(w, cfg, x[:, :, i:i + 1], cache, first_chunk=(i == 0), trace=trace) for i in range(z.shape[2])]
    return unpatchify(torch.cat(outs, dim=2))


def single_decode(w: Weights, cfg: VAE38Config, z: torch.Tensor) -> torch.Tensor:
    """WanVideoVAE.single_decode, VAE:1212-1215."""
    return model_decode(w, cfg, z).clamp_(-1, 1)


def build_1d_mask(length: int, left_bound: bool, right_bound: bool, border: int) -> torch.Tensor:
    """VAE:1081-1087."""
    x = torch.ones(length)
    if not left_bound:
        x[:border] = (torch.arange(border) + 1) / border
    if not right_bound:
        x[-border:] = torch.flip((torch.arange(border) + 1) / border, dims=(0,))
    return x


def tile_tasks(H: int, W: int, tile_size, tile_stride):
    """The (h, h_end, w, w_end) latent windows of tiled_decode, VAE:1108-1116."""
    (sh, sw), (th, tw) = tile_size, tile_stride
    tasks = []
    for h in range(0, H, th):
        if h - th >= 0 and h - th + sh >= H:
            continue
        for ww in range(0, W, tw):
            if ww - tw >= 0 and ww - tw + sw >= W:
                continue
            tasks.append((h, h + sh, ww, ww + sw))
    return tasks


def tiled_decode(w: Weights, cfg: VAE38Config, z: torch.Tensor, tile_size=(34, 34), tile_stride=(18, 16)) -> torch.Tensor:
    """WanVideoVAE.tiled_decode, VAE:1103-1152: overlapping latent windows decoded independently and blended with linear
    ramps of (size - stride) * 16 pixels on interior edges."""
    _, _, T, H, W = z.shape
    f = cfg.upsampling_factor
    out_t = T * 4 - 3
    weight = torch.zeros(1, 1, out_t, H * f, W * f, dtype=z.dtype)
    values = torch.zeros(1, 3, out_t, H * f, W * f, dtype=z.dtype)
    for h, h_, ww, w_ in tile_tasks(H, W, tile_size, tile_stride):
        part = model_decode(w, cfg, z[:, :, :, h:h_, ww:w_])
        ph, pw = part.shape[3], part.shape[4]
        mh = build_1d_mask(ph, h == 0, h_ >= H, (tile_size[0] - tile_stride[0]) * f)
        mw = build_1d_mask(pw, ww == 0, w_ >= W, (tile_size[1] - tile_stride[1]) * f)
        mask = torch.minimum(mh[:, None].expand(ph, pw), mw[None, :].expand(ph, pw)).view(1, 1, 1, ph, pw).to(z.dtype)
        values[:, :, :, h * f:h * f + ph, ww * f:ww * f + pw] += part * mask
        weight[:, :, :, h * f:h * f + ph, ww * f:ww * f + pw] += mask
    return (values / weight).clamp_(-1, 1)


def decode(w: Weights, cfg: VAE38Config, latents: torch.Tensor, tiled: bool = False, tile_size=(34, 34), tile_stride=(18, 16)):
    """WanVideoVAE.decode, VAE:1235-1248: latents [B, z_dim, T, h, w] -> videos [B, 3, 4T-3, 16h, 16w] in [-1, 1]."""
    vids = []
    for lat in latents:
        lat = lat.unsqueeze(0)
        vids.append((tiled_decode(w, cfg, lat, tile_size, tile_stride) if tiled else single_decode(w, cfg, lat)).squeeze(0))
    return torch.stack(vids)


# ============================================================================================
# encoder (first-frame conditioning: pipelines/wan_video.py:490-497 -> WanVideoVAE.encode, VAE:1218-1232)
# ============================================================================================
def enc_stage_plan(cfg: VAE38Config):
    """(in_dim, out_dim, n residual blocks, down_flag, temporal down) of the Down_ResidualBlocks, VAE:644-661."""
    dims = cfg.enc_dims
    plan = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        t_down = cfg.temperal_downsample[i] if i < len(cfg.temperal_downsample) else False
        plan.append((a, b, cfg.num_res_blocks, i != len(cfg.dim_mult) - 1, t_down))
    return plan


def enc_param_shapes(cfg: VAE38Config) -> Dict[str, tuple]:
    """State-dict keys / shapes of ``VideoVAE38_`` that encode touches: ``encoder`` and ``conv1`` (VAE:1291-1294, 620-676)."""
    d0, dl = cfg.enc_dims[0], cfg.enc_dims[-1]
    s = {"conv1.weight": (2 * cfg.z_dim, 2 * cfg.z_dim, 1, 1, 1), "conv1.bias": (2 * cfg.z_dim,),
         "encoder.conv1.weight": (d0, 12, 3, 3, 3), "encoder.conv1.bias": (d0,)}
    for i, (cin, cout, n, down, t_down) in enumerate(enc_stage_plan(cfg)):
        p = f"encoder.downsamples.{i}.downsamples."
        c = cin
        for j in range(n):
            s.update(_res_shapes(f"{p}{j}.", c, cout))
            c = cout
        if down:
            s[f"{p}{n}.resample.1.weight"] = (cout, cout, 3, 3)
            s[f"{p}{n}.resample.1.bias"] = (cout,)
            if t_down:
                s[f"{p}{n}.time_conv.weight"] = (cout, cout, 3, 1, 1)
                s[f"{p}{n}.time_conv.bias"] = (cout,)
    s.update(_res_shapes("encoder.middle.0.", dl, dl))
    s.update({"encoder.middle.1.norm.gamma": (dl, 1, 1), "encoder.middle.1.to_qkv.weight": (3 * dl, dl, 1, 1),
              "encoder.middle.1.to_qkv.bias": (3 * dl,), "encoder.middle.1.proj.weight": (dl, dl, 1, 1), "encoder.middle.1.proj.bias": (dl,)})
    s.update(_res_shapes("encoder.middle.2.", dl, dl))
    s.update({"encoder.head.0.gamma": (dl, 1, 1, 1), "encoder.head.2.weight": (2 * cfg.z_dim, dl, 3, 3, 3), "encoder.head.2.bias": (2 * cfg.z_dim,)})
    return s


def make_enc_weights(cfg: VAE38Config, seed: int = 0) -> Weights:
    out = {}
    for name, shape in enc_param_shapes(cfg).items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        if name.endswith("gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for v in shape[1:]:
                fan_in *= v
            t = torch.randn(shape, generator=g) * (fan_in ** -0.5)
        out[name] = t
    return out


def patchify(x, patch: int = 2):
    """'b c f (h q) (w r) -> b (c r q) f h w', VAE:199-211."""
    b, c, f, H, W = x.shape
    x = x.view(b, c, f, H // patch, patch, W // patch, patch)          # (h, q, w, r)
    return x.permute(0, 1, 6, 4, 2, 3, 5).reshape(b, c * patch * patch, f, H // patch, W // patch)


def avg_down3d(x, out_channels: int, factor_t: int, factor_s: int):
    """AvgDown3D.forward, VAE:363-395: (t, h, w) sub-positions fold into channels, groups of channels are averaged."""
    pad_t = (factor_t - x.shape[2] % factor_t) % factor_t
    x = F.pad(x, (0, 0, 0, 0, pad_t, 0))
    B, C, T, H, W = x.shape
    factor = factor_t * factor_s * factor_s
    x = x.view(B, C, T // factor_t, factor_t, H // factor_s, factor_s, W // factor_s, factor_s).permute(0, 1, 3, 5, 7, 2, 4, 6).contiguous()
    x = x.view(B, out_channels, C * factor // out_channels, T // factor_t, H // factor_s, W // factor_s)
    return x.mean(dim=2)


def resample_down(w: Weights, p: str, x, temporal: bool, cache: List, idx: List[int]):
    """Resample38.forward for 'downsample2d' / 'downsample3d', VAE:157-173: ZeroPad2d((0,1,0,1)) + 3x3 stride-2 Conv2d per
    frame; then, temporal: the first chunk only stores its frame, later chunks run the stride-2 time_conv on
    (last cached frame | x)."""
    b, c, t, hh, ww = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)
    y = F.conv2d(F.pad(y, (0, 1, 0, 1)), w[p + "resample.1.weight"], w[p + "resample.1.bias"], stride=2)
    x = y.reshape(b, t, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
    if temporal:
        i = idx[0]
        if cache[i] is None:
            cache[i] = x.clone()
        else:
            tail = x[:, :, -1:].clone()
            x = F.conv3d(torch.cat([cache[i][:, :, -1:], x], 2), w[p + "time_conv.weight"], w[p + "time_conv.bias"], stride=(2, 1, 1))
            cache[i] = tail
        idx[0] += 1
    return x


def encoder_chunk(w: Weights, cfg: VAE38Config, x, cache: List):
    """Encoder3d_38.forward on one chunk of frames, VAE:679-733."""
    idx = [0]
    tail = _cache_tail(x, cache[0])
    x = causal_conv3d(x, w["encoder.conv1.weight"], w["encoder.conv1.bias"], cache[0])
    cache[0] = tail
    idx[0] = 1
    for i, (cin, cout, n, down, t_down) in enumerate(enc_stage_plan(cfg)):            # Down_ResidualBlock.forward, VAE:469-474
        p = f"encoder.downsamples.{i}.downsamples."
        main = x
        for j in range(n):
            main = residual_block(w, f"{p}{j}.", main, cache, idx)
        if down:
            main = resample_down(w, f"{p}{n}.", main, t_down, cache, idx)
        x = main + avg_down3d(x, cout, 2 if t_down else 1, 2 if down else 1)
    x = residual_block(w, "encoder.middle.0.", x, cache, idx)
    x = attention_block(w, "encoder.middle.1.", x)
    x = residual_block(w, "encoder.middle.2.", x, cache, idx)
    x = F.silu(rms_norm(x, w["encoder.head.0.gamma"]))
    i = idx[0]
    tail = _cache_tail(x, cache[i])
    x = causal_conv3d(x, w["encoder.head.2.weight"], w["encoder.head.2.bias"], cache[i])
    cache[i] = tail
    return x


def count_enc_cache_slots(cfg: VAE38Config) -> int:
    """count_conv3d(encoder), VAE:943-948."""
    return sum(1 for k, s in enc_param_shapes(cfg).items() if k.startswith("encoder.") and k.endswith("weight") and len(s) == 5)


def model_encode(w: Weights, cfg: VAE38Config, video: torch.Tensor) -> torch.Tensor:
    """VideoVAE38_.encode, VAE:1298-1323: video [1, 3, 1+4k, H, W] in [-1, 1] -> normalised latent mean [1, z_dim, 1+k, H/16, W/16];
    the first frame alone, then chunks of 4 frames."""
    x = patchify(video)
    cache: List = [None] * count_enc_cache_slots(cfg)
    t = x.shape[2]
    outs = []
    for i in range(1 + (t - 1) // 4):
        chunk = x[:, :, :1] if i == 0 else x[:, :, 1 + 4 * (i - 1):1 + 4 * i]
        outs.append(encoder_chunk(w, cfg, chunk, cache))
    mu = causal_conv3d(torch.cat(outs, 2), w["conv1.weight"], w["conv1.bias"]).chunk(2, dim=1)[0]
    mean, inv_std = latent_scale(cfg)
    return (mu - mean.to(mu).view(1, -1, 1, 1, 1)) * inv_std.to(mu).view(1, -1, 1, 1, 1)


def tiled_encode(w: Weights, cfg: VAE38Config, video: torch.Tensor, tile_size, tile_stride) -> torch.Tensor:
    """WanVideoVAE.tiled_encode, VAE:1155-1203; tile_size / tile_stride in PIXELS (encode() multiplies the latent-unit arguments
    by 16 first, VAE:1224-1226)."""
    _, _, T, H, W = video.shape
    f = cfg.upsampling_factor
    out_t = (T + 3) // 4
    weight = torch.zeros(1, 1, out_t, H // f, W // f, dtype=video.dtype)
    values = torch.zeros(1, cfg.z_dim, out_t, H // f, W // f, dtype=video.dtype)
    for h, h_, ww, w_ in tile_tasks(H, W, tile_size, tile_stride):
        part = model_encode(w, cfg, video[:, :, :, h:h_, ww:w_])
        ph, pw = part.shape[3], part.shape[4]
        mh = build_1d_mask(ph, h == 0, h_ >= H, (tile_size[0] - tile_stride[0]) // f)
        mw = build_1d_mask(pw, ww == 0, w_ >= W, (tile_size[1] - tile_stride[1]) // f)
        mask = torch.minimum(mh[:, None].expand(ph, pw), mw[None, :].expand(ph, pw)).view(1, 1, 1, ph, pw).to(video.dtype)
        values[:, :, :, h // f:h // f + ph, ww // f:ww // f + pw] += part * mask
        weight[:, :, :, h // f:h // f + ph, ww // f:ww // f + pw] += mask
    return values / weight


def encode(w: Weights, cfg: VAE38Config, videos, tiled: bool = False, tile_size=(34, 34), tile_stride=(18, 16)) -> torch.Tensor:
    """WanVideoVAE.encode, VAE:1218-1232: list of videos [3, T, H, W] -> latents [B, z_dim, (T+3)//4, H/16, W/16]."""
    f = cfg.upsampling_factor
    outs = []
    for v in videos:
        v = v.unsqueeze(0)
        if tiled:
            outs.append(tiled_encode(w, cfg, v, (tile_size[0] * f, tile_size[1] * f), (tile_stride[0] * f, tile_stride[1] * f)).squeeze(0))
        else:
            outs.append(model_encode(w, cfg, v).squeeze(0))
    return torch.stack(outs)
