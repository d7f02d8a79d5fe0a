"""Wan2.2 VAE38 decoder on the sm_100a kernels.

Drop-in for the decode half of ``pipe.vae`` — the reference's ``WanVideoVAE38`` (models/wan_video_vae.py:1354-1382, VAE) as
``WanVideoPipeline.__call__`` uses it after the denoising loop (pipelines/wan_video.py:322-323):
``vae.decode(latents, device=..., tiled=..., tile_size=..., tile_stride=...)`` -> video ``[B, 3, 4T-3, 16h, 16w]`` in [-1, 1].
Same state-dict keys (``model.decoder.*``, ``model.conv2.*``), same chunk-by-chunk causal schedule with the two-frame feature
cache, same tiling and blending ramps.

How it runs here: every feature map is a channels-last, zero-bordered grid ``[T][H+2][W+2][Cp]`` (Cp = channels rounded up to
64) flattened to rows; every convolution (causal 3x3x3, 3x3 after the up-sampling, the 3x1x1 temporal one, 1x1) is ONE
``fgb_conv_taps_bf16`` launch — the tcgen05 GEMM reading shifted rows of the grid per kernel tap, with the cached frames of
``CausalConv3d`` simply stored in front of the new ones and the grid border re-zeroed by the epilogue — and the rest
(RMS_norm + SiLU, nearest-exact up-sampling with the temporal frame interleave, the DupUp3D shortcut, the attention block's
softmax, un-patchify + tile blending + clamp) are the one-pass kernels of ``csrc/vae.cu``.  No cuDNN, no CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import ops
from .ops import BF16, EPI_BIAS, EPI_RESIDUAL

# WanVideoVAE38.__init__ (VAE:1359-1376): per-channel statistics of the 48 latent channels
MEAN38 = [-0.2289, -0.0052, -0.1323, -0.2339, -0.2799, 0.0174, 0.1838, 0.1557, -0.1382, 0.0542, 0.2813, 0.0891, 0.1570, -0.0098, 0.0375,
          -0.1825, -0.2246, -0.1207, -0.0698, 0.5109, 0.2665, -0.2108, -0.2158, 0.2502, -0.2055, -0.0322, 0.1109, 0.1567, -0.0729, 0.0899,
          -0.2799, -0.1230, -0.0313, -0.1649, 0.0117, 0.0723, -0.2839, -0.2083, -0.0520, 0.3748, 0.0152, 0.1957, 0.1433, -0.2944, 0.3573,
          -0.0548, -0.1681, -0.0667]
STD38 = [0.4765, 1.0364, 0.4514, 1.1677, 0.5313, 0.4990, 0.4818, 0.5013, 0.8158, 1.0344, 0.5894, 1.0901, 0.6885, 0.6165, 0.8454, 0.4978,
         0.5759, 0.3523, 0.7135, 0.6804, 0.5833, 1.4146, 0.8986, 0.5659, 0.7069, 0.5338, 0.4889, 0.4917, 0.4069, 0.4999, 0.6866, 0.4093,
         0.5709, 0.6065, 0.6415, 0.4944, 0.5726, 1.2042, 0.5458, 1.6887, 0.3971, 1.0600, 0.3943, 0.5537, 0.5444, 0.4089, 0.7468, 0.7744]


@dataclass(frozen=True)
class VAE38Config:
    """VideoVAE38_.__init__ defaults (VAE:1271-1296); temperal_upsample is the reversed temperal_downsample (sic)."""
    z_dim: int = 48
    dec_dim: int = 256
    dim_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temperal_upsample: Tuple[bool, ...] = (True, True, False)

    @property
    def dims(self) -> List[int]:
        return [self.dec_dim * u for u in [self.dim_mult[-1]] + list(self.dim_mult[::-1])]


VAE38 = VAE38Config()


def param_shapes(cfg: VAE38Config) -> Dict[str, tuple]:
    """State-dict keys / shapes of ``VideoVAE38_`` that decode reads: ``conv2`` and ``decoder`` (VAE:1294-1296, 842-887)."""
    def res(p, cin, cout):
        s = {p + "residual.0.gamma": (cin, 1, 1, 1), p + "residual.2.weight": (cout, cin, 3, 3, 3), p + "residual.2.bias": (cout,),
             p + "residual.3.gamma": (cout, 1, 1, 1), p + "residual.6.weight": (cout, cout, 3, 3, 3), p + "residual.6.bias": (cout,)}
        if cin != cout:
            s.update({p + "shortcut.weight": (cout, cin, 1, 1, 1), p + "shortcut.bias": (cout,)})
        return s

    dims, z = cfg.dims, cfg.z_dim
    d0 = dims[0]
    out = {"conv2.weight": (z, z, 1, 1, 1), "conv2.bias": (z,), "decoder.conv1.weight": (d0, z, 3, 3, 3), "decoder.conv1.bias": (d0,)}
    out.update(res("decoder.middle.0.", d0, d0))
    out.update({"decoder.middle.1.norm.gamma": (d0, 1, 1), "decoder.middle.1.to_qkv.weight": (3 * d0, d0, 1, 1),
                "decoder.middle.1.to_qkv.bias": (3 * d0,), "decoder.middle.1.proj.weight": (d0, d0, 1, 1), "decoder.middle.1.proj.bias": (d0,)})
    out.update(res("decoder.middle.2.", d0, d0))
    n = cfg.num_res_blocks + 1
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        p = f"decoder.upsamples.{i}.upsamples."
        c = cin
        for j in range(n):
            out.update(res(f"{p}{j}.", c, cout))
            c = cout
        if i != len(cfg.dim_mult) - 1:
            out.update({f"{p}{n}.resample.1.weight": (cout, cout, 3, 3), f"{p}{n}.resample.1.bias": (cout,)})
            if i < len(cfg.temperal_upsample) and cfg.temperal_upsample[i]:
                out.update({f"{p}{n}.time_conv.weight": (2 * cout, cout, 3, 1, 1), f"{p}{n}.time_conv.bias": (2 * cout,)})
    out.update({"decoder.head.0.gamma": (dims[-1], 1, 1, 1), "decoder.head.2.weight": (12, dims[-1], 3, 3, 3), "decoder.head.2.bias": (12,)})
    return out


def _c64(c: int) -> int:
    return -(-c // 64) * 64


def tile_tasks(H: int, W: int, tile_size, tile_stride):
    """(h, h_end, w, w_end) latent windows of tiled_decode (VAE:1108-1116)."""
    (sh, sw), (th, tw) = tile_size, tile_stride
    tasks = []
    for h in range(0, H, th):
        if h - th >= 0 and h - th + sh >= H:
            continue
        for w in range(0, W, tw):
            if w - tw >= 0 and w - tw + sw >= W:
                continue
            tasks.append((h, h + sh, w, w + sw))
    return tasks


def assign_windows(tasks, H: int, W: int, world: int):
    """Windows of one video spread over `world` ranks: largest (clipped) area first, each to the least loaded rank — the
    windows are independent until the blend, so this is the multi-GPU split of the decode (no exchange but the final sum)."""
    area = lambda t: (min(t[1], H) - t[0]) * (min(t[3], W) - t[2])  # noqa: E731
    order = sorted(range(len(tasks)), key=lambda i: (-area(tasks[i]), i))
    load, mine = [0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        load[r] += area(tasks[i])
        mine[r].append(tasks[i])
    return mine


class _Conv:
    """One convolution as a tap GEMM: weight repacked to [parts*Cout_p, taps*Cin_p] (tap-major K), bias to [parts*Cout_p]."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, device, parts: int = 1):
        w = weight.detach().to(device=device, dtype=torch.float32)
        cout_total, cin = w.shape[0], w.shape[1]
        kernel = tuple(w.shape[2:])
        if len(kernel) == 2:
            kernel = (1,) + kernel
        self.kernel = kernel
        self.cin, self.cout, self.parts = cin, cout_total // parts, parts
        self.cin_p, self.cout_p = _c64(cin), _c64(self.cout)
        nt = kernel[0] * kernel[1] * kernel[2]
        wp = torch.zeros(parts, self.cout_p, nt, self.cin_p, dtype=torch.float32, device=device)
        wp[:, :self.cout, :, :cin] = w.reshape(parts, self.cout, cin, nt).permute(0, 1, 3, 2)
        self.w = wp.reshape(parts * self.cout_p, nt * self.cin_p).to(BF16).contiguous()
        bp = torch.zeros(parts, self.cout_p, dtype=torch.float32, device=device)
        bp[:, :self.cout] = bias.detach().to(device=device, dtype=torch.float32).view(parts, self.cout)
        self.b = bp.reshape(-1).to(BF16).contiguous()
        self.hist = kernel[0] - 1                     # cached frames in front of the new ones (CausalConv3d, VAE:44-52)
        self.n = parts * self.cout_p
        self.inp: Optional[torch.Tensor] = None       # [(hist + Tmax) * P, cin_p] rows, allocated per tile geometry

    def offsets(self, hp: int, wp: int) -> List[int]:
        kt, kh, kw = self.kernel
        return [((it - (kt - 1)) * hp + (iy - kh // 2)) * wp + (ix - kw // 2) for it in range(kt) for iy in range(kh) for ix in range(kw)]


class _Res:
    __slots__ = ("g1", "c1", "g2", "c2", "short", "cin", "cout")


class VAE38Decoder:
    def __init__(self, cfg: VAE38Config = VAE38, device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        ops.context(self.device)                      # raises off-GPU: there is no CPU path
        self.loaded = False
        self.upsampling_factor = 2 ** (len(cfg.dim_mult) - 1) * 2          # 16 (VAE:1380)
        self.z_dim = cfg.z_dim
        self.kernel_launches = 0
        self.trace: Optional[Callable] = None         # debug / tests: trace(name, rows, T, h, w, channels)
        self._geom = None
        n = min(cfg.z_dim, len(MEAN38))
        self.mean = torch.tensor(MEAN38[:n], dtype=torch.float32, device=self.device)
        self.inv_std = (1.0 / torch.tensor(STD38[:n], dtype=torch.float32)).to(self.device)

    # ------------------------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------------------------
    def _stage_plan(self):
        cfg, dims = self.cfg, self.cfg.dims
        plan = []
        for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
            t_up = cfg.temperal_upsample[i] if i < len(cfg.temperal_upsample) else False
            plan.append((a, b, cfg.num_res_blocks + 1, i != len(cfg.dim_mult) - 1, t_up))
        return plan

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Keys of ``WanVideoVAE38`` (``model.decoder.*``, ``model.conv2.*``) or of ``VideoVAE38_`` (no ``model.`` prefix);
        encoder keys are ignored."""
        if any(k.startswith("model.") for k in sd):
            sd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
        dev = self.device
        for k, shape in param_shapes(self.cfg).items():
            if k not in sd:
                raise KeyError(f"VAE38 state dict has no '{k}'")
            if tuple(sd[k].shape) != shape:
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {shape} of the configuration")

        def need(k):
            return sd[k]

        def conv(p, parts=1):
            return _Conv(need(p + "weight"), need(p + "bias"), dev, parts)

        def gamma(k, c):
            g = torch.zeros(_c64(c), dtype=torch.float32, device=dev)
            v = need(k).detach().to(device=dev, dtype=torch.float32).reshape(-1)
            if v.numel() != c:
                raise ValueError(f"{k}: {v.numel()} channels, expected {c}")
            g[:c] = v
            return g.to(BF16)

        def res(p, cin, cout):
            r = _Res()
            r.cin, r.cout = cin, cout
            r.g1, r.c1 = gamma(p + "residual.0.gamma", cin), conv(p + "residual.2.")
            r.g2, r.c2 = gamma(p + "residual.3.gamma", cout), conv(p + "residual.6.")
            r.short = conv(p + "shortcut.") if cin != cout else None
            if r.c1.cin != cin or r.c1.cout != cout or r.c2.cout != cout:
                raise ValueError(f"{p}: convolution shapes do not match the configuration ({cin} -> {cout})")
            return r

        cfg = self.cfg
        d0 = cfg.dims[0]
        self.conv2 = conv("conv2.")
        self.conv1 = conv("decoder.conv1.")
        if self.conv1.cin != cfg.z_dim or self.conv1.cout != d0:
            raise ValueError(f"decoder.conv1 is {self.conv1.cin} -> {self.conv1.cout}, configuration says {cfg.z_dim} -> {d0}")
        self.mid0 = res("decoder.middle.0.", d0, d0)
        self.attn_g = gamma("decoder.middle.1.norm.gamma", d0)
        self.attn_qkv = conv("decoder.middle.1.to_qkv.", parts=3)
        self.attn_proj = conv("decoder.middle.1.proj.")
        self.mid2 = res("decoder.middle.2.", d0, d0)
        self.stages = []
        for i, (cin, cout, n, up, t_up) in enumerate(self._stage_plan()):
            p = f"decoder.upsamples.{i}.upsamples."
            blocks, c = [], cin
            for j in range(n):
                blocks.append(res(f"{p}{j}.", c, cout))
                c = cout
            st = dict(cin=cin, cout=cout, blocks=blocks, up=up, t_up=t_up, conv2d=None, time_conv=None)
            if up:
                st["conv2d"] = conv(f"{p}{n}.resample.1.")
                if t_up:
                    st["time_conv"] = conv(f"{p}{n}.time_conv.", parts=2)
            self.stages.append(st)
        self.head_g = gamma("decoder.head.0.gamma", cfg.dims[-1])
        self.head = conv("decoder.head.2.")
        self._geom = None
        self.loaded = True

    # ------------------------------------------------------------------------------------------------------------
    # one convolution
    # ------------------------------------------------------------------------------------------------------------
    def _alloc(self, rows: int, cols: int) -> torch.Tensor:
        """Zeroed rows: for grids whose border / slack rows no kernel writes."""
        return torch.zeros(rows, cols, dtype=BF16, device=self.device)

    def _empty(self, rows: int, cols: int) -> torch.Tensor:
        """Rows that the next kernel writes completely (convolution, norm and GEMM outputs cover every row and column)."""
        return torch.empty(rows, cols, dtype=BF16, device=self.device)

    def _input(self, c: _Conv, T: int, h: int, w: int) -> torch.Tensor:
        """The rows of `c`'s input buffer that the T new frames go to (the cached frames sit in front of them)."""
        P = (h + 2) * (w + 2)
        need = (c.hist + T) * P
        if c.inp is None or c.inp.shape[1] != c.cin_p:
            c.inp = self._alloc(need, c.cin_p)
        elif c.inp.shape[0] < need:      # later chunks carry more frames than the first: grow, keeping the cached frames
            grown = self._alloc(need, c.cin_p)
            grown[:c.hist * P].copy_(c.inp[:c.hist * P])
            c.inp = grown
        return c.inp[c.hist * P:(c.hist + T) * P]

    def _conv(self, c: _Conv, T: int, h: int, w: int, out: torch.Tensor, epilogue=EPI_BIAS, src: Optional[torch.Tensor] = None, mask=True):
        """out[T*P, n] = conv(c) of the T new frames. src = None: the input was written into c's own buffer via _input() and
        the cached frames are rolled afterwards; src given: a convolution without temporal extent reading `src` directly."""
        hp, wp = h + 2, w + 2
        P = hp * wp
        if src is None:
            x = c.inp[:(c.hist + T) * P]
            ops.conv_taps(x, c.hist * P, c.w, c.b, out[:T * P], c.offsets(hp, wp), (hp, wp) if mask else (0, 0), epilogue)
            if c.hist:   # the last `hist` frames become the cache of the next chunk (VAE:288-301)
                tail = c.inp[T * P:(T + c.hist) * P].clone()
                c.inp[:c.hist * P].copy_(tail)
        else:
            if c.hist:
                raise ValueError("a causal convolution needs its own input buffer")
            ops.conv_taps(src[:T * P], 0, c.w, c.b, out[:T * P], c.offsets(hp, wp), (hp, wp) if mask else (0, 0), epilogue)
        self.kernel_launches += 1
        return out

    def _tr(self, name, rows, T, h, w, channels):
        if self.trace is not None:
            self.trace(name, rows[:T * (h + 2) * (w + 2)], T, h, w, channels)

    # ------------------------------------------------------------------------------------------------------------
    # blocks
    # ------------------------------------------------------------------------------------------------------------
    def _res(self, r: _Res, x: torch.Tensor, T: int, h: int, w: int, tag: str, keep_input: bool = True) -> torch.Tensor:
        """ResidualBlock.forward (VAE:283-301). keep_input = False: nobody reads `x` afterwards, so an identity shortcut
        accumulates the second convolution straight into it instead of into a copy."""
        P = (h + 2) * (w + 2)
        rows = T * P
        if r.short is not None:
            out = self._empty(rows, r.c2.cout_p)
            self._conv(r.short, T, h, w, out, EPI_BIAS, src=x)
        elif keep_input:
            out = x[:rows].clone()
        else:
            out = x[:rows]
        ops.vae_norm_silu(x[:rows], self._input(r.c1, T, h, w), r.cin, r.g1)
        tmp = self._empty(rows, r.c1.cout_p)
        self._conv(r.c1, T, h, w, tmp)
        ops.vae_norm_silu(tmp, self._input(r.c2, T, h, w), r.cout, r.g2)
        self._conv(r.c2, T, h, w, out, EPI_RESIDUAL)
        self.kernel_launches += 2
        self._tr(tag, out, T, h, w, r.cout)
        return out

    def _attention(self, x: torch.Tensor, T: int, h: int, w: int) -> torch.Tensor:
        """AttentionBlock.forward (VAE:321-342): one head of width C over the positions of each frame; the grid border is
        excluded as keys and re-zeroed in the output."""
        C = getattr(self, "_attn_dim", self.cfg.dims[0])      # the encoder (vae_encode.py) runs the same block at its own width
        cp = _c64(C)
        hp, wp = h + 2, w + 2
        P = hp * wp
        n8 = -(-P // 8) * 8
        rows = T * P
        nbuf = self._empty(rows, cp)
        ops.vae_norm_silu(x[:rows], nbuf, C, self.attn_g, silu=False)
        qkv = self._alloc(rows + 8, 3 * cp)
        self._conv(self.attn_qkv, T, h, w, qkv, EPI_BIAS, src=nbuf, mask=False)
        o = self._empty(rows, cp)
        s = torch.empty(P, n8, dtype=BF16, device=self.device)
        for f in range(T):
            q = qkv[f * P:(f + 1) * P, :cp]
            k = qkv[f * P:f * P + n8, cp:2 * cp]
            v = qkv[f * P:f * P + n8, 2 * cp:]
            ops.gemm(q, k, None, s)
            ops.vae_attn_softmax(s, n8, hp, wp, 1.0 / math.sqrt(C))
            ops.gemm_dgrad(s, v, o[f * P:(f + 1) * P])
        out = x[:rows].clone()
        self._conv(self.attn_proj, T, h, w, out, EPI_RESIDUAL, src=o)
        self.kernel_launches += 1 + 3 * T
        self._tr("attn", out, T, h, w, C)
        return out

    def _stage(self, i: int, st, x: torch.Tensor, T: int, h: int, w: int, first_chunk: bool):
        """Up_ResidualBlock.forward (VAE:506-514) with Resample38 'upsample2d' / 'upsample3d' (VAE:120-160)."""
        main = x
        for j, r in enumerate(st["blocks"]):   # the stage input `x` feeds the DupUp3D shortcut later: block 0 must not overwrite it
            main = self._res(r, main, T, h, w, f"s{i}.b{j}", keep_input=(j == 0 and st["up"]))
        if not st["up"]:
            return main, T, h, w
        cout, cp = st["cout"], _c64(st["cout"])
        P = (h + 2) * (w + 2)
        T2 = T
        if st["t_up"] and not first_chunk:      # the first chunk is not doubled in time ('Rep', VAE:124-127)
            tc = st["time_conv"]
            self._input(tc, T, h, w).copy_(main[:T * P])
            both = self._empty(T * P, 2 * cp)
            self._conv(tc, T, h, w, both)
            T2 = 2 * T
            up = self._alloc(T2 * (2 * h + 2) * (2 * w + 2), cp)
            ops.vae_upsample2x(both, up, cp, T2, h, w, halves=2)
        else:
            up = self._alloc(T2 * (2 * h + 2) * (2 * w + 2), cp)
            ops.vae_upsample2x(main, up, cp, T2, h, w, halves=1)
        out = self._empty(T2 * (2 * h + 2) * (2 * w + 2), cp)
        self._conv(st["conv2d"], T2, 2 * h, 2 * w, out, EPI_BIAS, src=up)
        self._tr(f"s{i}.resample", out, T2, 2 * h, 2 * w, cout)
        ops.vae_dup_up_add(x, out, st["cin"], cout, 2 if st["t_up"] else 1, first_chunk, T2, h, w)
        self.kernel_launches += 2
        self._tr(f"s{i}.out", out, T2, 2 * h, 2 * w, cout)
        return out, T2, 2 * h, 2 * w

    def _reset(self, h: int, w: int) -> None:
        """clear_cache() (VAE:1048-1055): forget the cached frames; buffers of another tile geometry are dropped."""
        convs = [self.conv1, self.head, self.mid0.c1, self.mid0.c2, self.mid2.c1, self.mid2.c2]
        for st in self.stages:
            for r in st["blocks"]:
                convs += [r.c1, r.c2]
            if st["time_conv"] is not None:
                convs.append(st["time_conv"])
        for c in convs:
            if self._geom != (h, w):
                c.inp = None
            elif c.inp is not None:
                c.inp.zero_()
        self._geom = (h, w)

    def _decode_window(self, z: torch.Tensor, values: torch.Tensor, weight: Optional[torch.Tensor], y0: int, x0: int, bounds, border) -> None:
        """VideoVAE38_.decode (VAE:1326-1351) of one latent window z [C, T, h, w] into the video at pixel (y0, x0)."""
        cfg = self.cfg
        C, T, h, w = z.shape
        self._reset(h, w)
        P = (h + 2) * (w + 2)
        zp = _c64(C)
        zg = self._alloc(T * P, zp)
        ops.vae_latent_rows(z, self.mean, self.inv_std, zg, zp)
        x_all = self._empty(T * P, self.conv2.cout_p)
        self._conv(self.conv2, T, h, w, x_all, EPI_BIAS, src=zg)
        self.kernel_launches += 1
        for i in range(T):
            first = i == 0
            self._input(self.conv1, 1, h, w).copy_(x_all[i * P:(i + 1) * P])
            x = self._empty(P, self.conv1.cout_p)
            self._conv(self.conv1, 1, h, w, x)
            self._tr("conv1", x, 1, h, w, cfg.dims[0])
            x = self._res(self.mid0, x, 1, h, w, "mid0", keep_input=False)
            x = self._attention(x, 1, h, w)
            x = self._res(self.mid2, x, 1, h, w, "mid2", keep_input=False)
            tc, hc, wc = 1, h, w
            for si, st in enumerate(self.stages):
                x, tc, hc, wc = self._stage(si, st, x, tc, hc, wc, first)
            rows = tc * (hc + 2) * (wc + 2)
            ops.vae_norm_silu(x[:rows], self._input(self.head, tc, hc, wc), cfg.dims[-1], self.head_g)
            hb = self._empty(rows, self.head.cout_p)
            self._conv(self.head, tc, hc, wc, hb)
            self._tr("head", hb, tc, hc, wc, 12)
            t0 = 0 if first else 1 + 4 * (i - 1)
            ops.vae_unpatchify(hb, tc, hc, wc, values, weight, t0, y0, x0, bounds, border)
            self.kernel_launches += 2

    # ------------------------------------------------------------------------------------------------------------
    # public: WanVideoVAE.decode (VAE:1235-1248)
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, hidden_states, device=None, tiled: bool = False, tile_size=(34, 34), tile_stride=(18, 16), group=None) -> torch.Tensor:
        """latents [B, z_dim, T, h, w] (tensor or list of [z_dim, T, h, w]) -> videos [B, 3, 4T-3, 16h, 16w] in [-1, 1], on the
        GPU in the latents' dtype.  `device` is accepted for signature compatibility (the decoder lives on its own device).
        group (a torch.distributed process group, tiled only): every rank holds the same latents and decodes its share of the
        windows (assign_windows); one sum all-reduce of the blended video and its weights gives every rank the full result."""
        if not self.loaded:
            raise RuntimeError("VAE38Decoder.decode called before load_state_dict")
        world, rank = 1, 0
        if group is not None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        vids = []
        f = self.upsampling_factor
        for lat in hidden_states:
            if lat.dim() != 4 or lat.shape[0] != self.cfg.z_dim:
                raise ValueError(f"each latent must be [{self.cfg.z_dim}, T, h, w], got {tuple(lat.shape)}")
            dtype = lat.dtype
            z = lat.to(device=self.device, dtype=BF16).contiguous()
            _, T, H, W = z.shape
            out_t = 4 * T - 3
            values = torch.zeros(3, out_t, H * f, W * f, dtype=torch.float32, device=self.device)
            if tiled:
                if tile_size[0] <= tile_stride[0] or tile_size[1] <= tile_stride[1]:
                    raise ValueError("tile_size must exceed tile_stride in both directions (the overlap carries the blending ramp)")
                weight = torch.zeros(out_t, H * f, W * f, dtype=torch.float32, device=self.device)
                border = ((tile_size[0] - tile_stride[0]) * f, (tile_size[1] - tile_stride[1]) * f)
                tasks = tile_tasks(H, W, tile_size, tile_stride)
                for h0, h1, w0, w1 in (assign_windows(tasks, H, W, world)[rank] if world > 1 else tasks):
                    win = z[:, :, h0:h1, w0:w1].contiguous()
                    self._decode_window(win, values, weight, h0 * f, w0 * f, (h0 == 0, h1 >= H, w0 == 0, w1 >= W), border)
                if world > 1:
                    dist.all_reduce(values, group=group)
                    dist.all_reduce(weight, group=group)
                ops.vae_blend_finish(values, weight)
                self.kernel_launches += 1
            else:
                self._decode_window(z, values, None, 0, 0, (True, True, True, True), (1, 1))
            vids.append(values.to(dtype))
        return torch.stack(vids)


def install(pipe) -> VAE38Decoder:
    """Route ``pipe.vae.decode`` (PIPE:322-323) through a VAE38Decoder holding the weights of the loaded reference VAE; its
    ``encode`` (first-frame conditioning, PIPE:495) is left on the reference module."""
    ref = getattr(pipe, "vae", None)
    if ref is None or not hasattr(ref, "model") or getattr(ref, "z_dim", None) != 48:
        raise ValueError("pipe.vae must be a loaded WanVideoVAE38")
    dec = VAE38Decoder(VAE38, getattr(pipe, "device", "cuda"))
    dec.load_state_dict(ref.state_dict())
    ref.decode = dec.decode
    return dec
