"""Wan2.2 VAE38 encoder on the sm_100a kernels — the encode half of ``pipe.vae``.

Reference: ``WanVideoVAE38.encode`` (models/wan_video_vae.py:1218-1232, VAE) -> ``VideoVAE38_.encode`` (VAE:1298-1323) ->
``Encoder3d_38`` (VAE:620-733): patchify, the first frame alone and then chunks of 4 frames through the causal encoder with
its feature cache, the 1x1x1 ``conv1``, the mean half of the channels, latent normalisation; used for the first-frame
conditioning of every generation (pipelines/wan_video.py:490-497) and for the training videos.

Same grid layout and the same convolution entry as the decoder (``vae.py``).  What is new on this side:
  * the stride-2 3x3 convolution behind ``ZeroPad2d((0, 1, 0, 1))`` runs on a space-to-depth copy of its input: 2x2 taps over
    4*Cp channels, kernel rows / columns 2*tap + phase, the fourth one zero — a stride-1 tap GEMM again;
  * the stride-2 temporal convolution is one tap-GEMM launch per output frame over the frames (2j-1, 2j, 2j+1) of
    (last cached frame | chunk);
  * ``AvgDown3D`` (the parameter-free shortcut) is a gather-mean kernel, with the zero frame the reference pads in front of the
    single-frame first chunk.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .ops import BF16, EPI_BIAS
from .vae import VAE38, VAE38Config, VAE38Decoder, _c64, _Conv, _Res, tile_tasks

ENC_DIM = 160                                   # `dim` of WanVideoVAE38 (VAE:1356)
TEMPERAL_DOWNSAMPLE = (False, True, True)       # VAE:1278


def enc_dims(cfg: VAE38Config, enc_dim: int = ENC_DIM) -> List[int]:
    return [enc_dim * u for u in [1] + list(cfg.dim_mult)]                           # VAE:638


def enc_stage_plan(cfg: VAE38Config, enc_dim: int = ENC_DIM):
    """(in, out, n residual blocks, down_flag, temporal down) of the Down_ResidualBlocks (VAE:644-661)."""
    dims = enc_dims(cfg, enc_dim)
    return [(a, b, cfg.num_res_blocks, i != len(cfg.dim_mult) - 1, TEMPERAL_DOWNSAMPLE[i] if i < len(TEMPERAL_DOWNSAMPLE) else False)
            for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))]


def enc_param_shapes(cfg: VAE38Config, enc_dim: int = ENC_DIM) -> Dict[str, tuple]:
    """State-dict keys / shapes of ``VideoVAE38_`` that encode reads: ``encoder`` and ``conv1`` (VAE:1291-1294, 620-676)."""
    def res(p, cin, cout):
        s = {p + "residual.0.gamma": (cin, 1, 1, 1), p + "residual.2.weight": (cout, cin, 3, 3, 3), p + "residual.2.bias": (cout,),
             p + "residual.3.gamma": (cout, 1, 1, 1), p + "residual.6.weight": (cout, cout, 3, 3, 3), p + "residual.6.bias": (cout,)}
        if cin != cout:
            s.update({p + "shortcut.weight": (cout, cin, 1, 1, 1), p + "shortcut.bias": (cout,)})
        return s

    z2 = 2 * cfg.z_dim
    dims = enc_dims(cfg, enc_dim)
    out = {"conv1.weight": (z2, z2, 1, 1, 1), "conv1.bias": (z2,), "encoder.conv1.weight": (dims[0], 12, 3, 3, 3), "encoder.conv1.bias": (dims[0],)}
    for i, (cin, cout, n, down, t_down) in enumerate(enc_stage_plan(cfg, enc_dim)):
        p = f"encoder.downsamples.{i}.downsamples."
        c = cin
        for j in range(n):
            out.update(res(f"{p}{j}.", c, cout))
            c = cout
        if down:
            out.update({f"{p}{n}.resample.1.weight": (cout, cout, 3, 3), f"{p}{n}.resample.1.bias": (cout,)})
            if t_down:
                out.update({f"{p}{n}.time_conv.weight": (cout, cout, 3, 1, 1), f"{p}{n}.time_conv.bias": (cout,)})
    dl = dims[-1]
    out.update(res("encoder.middle.0.", dl, dl))
    out.update({"encoder.middle.1.norm.gamma": (dl, 1, 1), "encoder.middle.1.to_qkv.weight": (3 * dl, dl, 1, 1),
                "encoder.middle.1.to_qkv.bias": (3 * dl,), "encoder.middle.1.proj.weight": (dl, dl, 1, 1), "encoder.middle.1.proj.bias": (dl,)})
    out.update(res("encoder.middle.2.", dl, dl))
    out.update({"encoder.head.0.gamma": (dl, 1, 1, 1), "encoder.head.2.weight": (z2, dl, 3, 3, 3), "encoder.head.2.bias": (z2,)})
    return out


def stride2_weight(weight: torch.Tensor, cin_p: int, cout_p: int) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [cout_p, 4 taps * 4 phases * cin_p] for the space-to-depth form of the stride-2 convolution:
    tap (ty, tx), phase (py, px) carries kernel element (2*ty + py, 2*tx + px); the fourth row / column does not exist (zero)."""
    cout, cin = weight.shape[:2]
    w = weight.detach().to(torch.float32)
    out = torch.zeros(cout_p, 2, 2, 2, 2, cin_p, dtype=torch.float32, device=w.device)      # [co, ty, tx, py, px, ci]
    for ty in range(2):
        for tx in range(2):
            for py in range(2):
                for px in range(2):
                    dy, dx = 2 * ty + py, 2 * tx + px
                    if dy < 3 and dx < 3:
                        out[:cout, ty, tx, py, px, :cin] = w[:, :, dy, dx]
    return out.reshape(cout_p, 16 * cin_p)


class _ConvS2(_Conv):
    """The stride-2 3x3 convolution of 'downsample2d/3d' as a 2x2-tap convolution over space-to-depth rows."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, device):   # noqa: D401 — does not call _Conv.__init__ (other packing)
        cout, cin = weight.shape[:2]
        self.kernel = (1, 2, 2)
        self.cin, self.cout, self.parts = cin, cout, 1
        self.cout_p = _c64(cout)
        self.cin_p = 4 * _c64(cin)                                            # channels of one space-to-depth row
        self.w = stride2_weight(weight.to(device), _c64(cin), self.cout_p).to(BF16).contiguous()
        b = torch.zeros(self.cout_p, dtype=torch.float32, device=device)
        b[:cout] = bias.detach().to(device=device, dtype=torch.float32)
        self.b = b.to(BF16)
        self.hist, self.n, self.inp = 0, self.cout_p, None

    def offsets(self, hp: int, wp: int) -> List[int]:
        return [ty * wp + tx for ty in range(2) for tx in range(2)]            # forward taps: the zero pad is bottom / right


class VAE38Encoder(VAE38Decoder):
    """Reuses the decoder's convolution / residual-block / attention plumbing (``_conv``, ``_res``, ``_attention``, buffers)."""

    def __init__(self, cfg: VAE38Config = VAE38, device="cuda", enc_dim: int = ENC_DIM):
        super().__init__(cfg, device)
        self.enc_dim = enc_dim

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        if any(k.startswith("model.") for k in sd):
            sd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
        cfg, dev = self.cfg, self.device
        for k, shape in enc_param_shapes(cfg, self.enc_dim).items():
            if k not in sd:
                raise KeyError(f"VAE38 state dict has no '{k}'")
            if tuple(sd[k].shape) != shape:
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {shape} of the configuration")

        def conv(p, parts=1):
            return _Conv(sd[p + "weight"], sd[p + "bias"], dev, parts)

        def gamma(k, c):
            g = torch.zeros(_c64(c), dtype=torch.float32, device=dev)
            g[:c] = sd[k].detach().to(device=dev, dtype=torch.float32).reshape(-1)
            return g.to(BF16)

        def res(p, cin, cout):
            r = _Res()
            r.cin, r.cout = cin, cout
            r.g1, r.c1 = gamma(p + "residual.0.gamma", cin), conv(p + "residual.2.")
            r.g2, r.c2 = gamma(p + "residual.3.gamma", cout), conv(p + "residual.6.")
            r.short = conv(p + "shortcut.") if cin != cout else None
            return r

        dims = enc_dims(cfg, self.enc_dim)
        dl = dims[-1]
        self.enc_conv1 = conv("encoder.conv1.")
        self.enc_stages = []
        for i, (cin, cout, n, down, t_down) in enumerate(enc_stage_plan(cfg, self.enc_dim)):
            p = f"encoder.downsamples.{i}.downsamples."
            blocks, c = [], cin
            for j in range(n):
                blocks.append(res(f"{p}{j}.", c, cout))
                c = cout
            st = dict(cin=cin, cout=cout, blocks=blocks, down=down, t_down=t_down, conv2d=None, time_conv=None, seen=False)
            if down:
                st["conv2d"] = _ConvS2(sd[f"{p}{n}.resample.1.weight"], sd[f"{p}{n}.resample.1.bias"], dev)
                if t_down:
                    st["time_conv"] = conv(f"{p}{n}.time_conv.")
            self.enc_stages.append(st)
        self.mid0, self.mid2 = res("encoder.middle.0.", dl, dl), res("encoder.middle.2.", dl, dl)
        self.attn_g = gamma("encoder.middle.1.norm.gamma", dl)
        self.attn_qkv = conv("encoder.middle.1.to_qkv.", parts=3)
        self.attn_proj = conv("encoder.middle.1.proj.")
        self._attn_dim = dl
        self.head_g = gamma("encoder.head.0.gamma", dl)
        self.head = conv("encoder.head.2.")
        self.out_conv = conv("conv1.")
        self._geom = None
        self.loaded = True

    # ------------------------------------------------------------------------------------------------------------
    def _reset(self, h: int, w: int) -> None:
        convs = [self.enc_conv1, self.head, self.mid0.c1, self.mid0.c2, self.mid2.c1, self.mid2.c2]
        for st in self.enc_stages:
            st["seen"] = False
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

    def _down_stage(self, i: int, st, x: torch.Tensor, T: int, h: int, w: int):
        """Down_ResidualBlock.forward (VAE:469-474) with Resample38 'downsample2d' / 'downsample3d' (VAE:157-173)."""
        main = x
        for j, r in enumerate(st["blocks"]):          # the stage input feeds the AvgDown3D shortcut: block 0 must not overwrite it
            main = self._res(r, main, T, h, w, f"d{i}.b{j}", keep_input=(j == 0))
        cout, cp = st["cout"], _c64(st["cout"])
        T2, h2, w2 = T, h, w
        if st["down"]:
            h2, w2 = h // 2, w // 2
            P2 = (h2 + 2) * (w2 + 2)
            s2d = self._alloc(T * P2, 4 * cp)
            ops.vae_space_to_depth(main, s2d, cp, T, h, w)
            down = self._empty(T * P2, cp)
            self._conv(st["conv2d"], T, h2, w2, down, EPI_BIAS, src=s2d)
            self.kernel_launches += 1
            main = down
            if st["t_down"]:
                tc = st["time_conv"]
                self._input(tc, T, h2, w2).copy_(down[:T * P2])
                if st["seen"]:                        # later chunks: stride 2 over (last cached frame | chunk) (VAE:166-171)
                    T2 = T // 2
                    main = self._empty(T2 * P2, cp)
                    for j in range(T2):
                        ops.conv_taps(tc.inp[:(tc.hist + T) * P2], (tc.hist + 2 * j + 1) * P2, tc.w, tc.b, main[j * P2:(j + 1) * P2],
                                      tc.offsets(h2 + 2, w2 + 2), (h2 + 2, w2 + 2), EPI_BIAS)
                    self.kernel_launches += T2
                st["seen"] = True                     # the first chunk only stores its frame (VAE:162-164)
                tail = tc.inp[T * P2:(T + tc.hist) * P2].clone()
                tc.inp[:tc.hist * P2].copy_(tail)
        ft, fs = (2 if st["t_down"] else 1), (2 if st["down"] else 1)
        ops.vae_avg_down_add(x, main, st["cin"], cout, ft, fs, (ft - T % ft) % ft, T2, h2, w2)
        self.kernel_launches += 1
        self._tr(f"d{i}.out", main, T2, h2, w2, cout)
        return main, T2, h2, w2

    def _encode_window(self, video: torch.Tensor, values: torch.Tensor, weight: Optional[torch.Tensor], y0: int, x0: int, bounds, border):
        """VideoVAE38_.encode (VAE:1298-1323) of one window [3, T, H, W] (T = 1 + 4k) into the latent at (y0, x0)."""
        cfg = self.cfg
        _, T, H, W = video.shape
        if H % 16 or W % 16 or (T - 1) % 4:
            raise ValueError(f"video window must be [3, 1+4k, 16a, 16b], got {tuple(video.shape)}")
        h, w = H // 2, W // 2
        self._reset(h, w)
        P = (h + 2) * (w + 2)
        grid = self._alloc(T * P, 64)
        ops.vae_patchify_rows(video, grid, 64)
        self.kernel_launches += 1
        n_chunks = 1 + (T - 1) // 4
        hl, wl = H // 16, W // 16
        Pl = (hl + 2) * (wl + 2)
        heads = self._empty(n_chunks * Pl, self.head.cout_p)
        for i in range(n_chunks):
            f0, tc = (0, 1) if i == 0 else (1 + 4 * (i - 1), 4)
            self._input(self.enc_conv1, tc, h, w).copy_(grid[f0 * P:(f0 + tc) * P])
            x = self._empty(tc * P, self.enc_conv1.cout_p)
            self._conv(self.enc_conv1, tc, h, w, x)
            hc, wc = h, w
            for si, st in enumerate(self.enc_stages):
                x, tc, hc, wc = self._down_stage(si, st, x, tc, hc, wc)
            x = self._res(self.mid0, x, tc, hc, wc, "mid0", keep_input=False)
            x = self._attention(x, tc, hc, wc)
            x = self._res(self.mid2, x, tc, hc, wc, "mid2", keep_input=False)
            rows = tc * Pl
            ops.vae_norm_silu(x[:rows], self._input(self.head, tc, hc, wc), enc_dims(cfg, self.enc_dim)[-1], self.head_g)
            self._conv(self.head, tc, hc, wc, heads[i * Pl:(i + 1) * Pl])
            self.kernel_launches += 1
        mu = self._empty(n_chunks * Pl, self.out_conv.cout_p)
        self._conv(self.out_conv, n_chunks, hl, wl, mu, EPI_BIAS, src=heads)
        self._tr("mu", mu, n_chunks, hl, wl, 2 * cfg.z_dim)
        ops.vae_latent_out(mu, n_chunks, hl, wl, self.mean, self.inv_std, values, weight, 0, y0, x0, bounds, border)
        self.kernel_launches += 1

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, videos, device=None, tiled: bool = False, tile_size=(34, 34), tile_stride=(18, 16)) -> torch.Tensor:
        """WanVideoVAE.encode (VAE:1218-1232): list / batch of videos [3, T, H, W] in [-1, 1] -> latents
        [B, z_dim, (T+3)//4, H/16, W/16] in the videos' dtype, on the GPU."""
        if not self.loaded:
            raise RuntimeError("VAE38Encoder.encode called before load_state_dict")
        f = self.upsampling_factor
        outs = []
        for vid in videos:
            if vid.dim() != 4 or vid.shape[0] != 3:
                raise ValueError(f"each video must be [3, T, H, W], got {tuple(vid.shape)}")
            dtype = vid.dtype
            v = vid.to(device=self.device, dtype=BF16).contiguous()
            _, T, H, W = v.shape
            values = torch.zeros(self.cfg.z_dim, (T + 3) // 4, H // f, W // f, dtype=torch.float32, device=self.device)
            if tiled:
                if tile_size[0] <= tile_stride[0] or tile_size[1] <= tile_stride[1]:
                    raise ValueError("tile_size must exceed tile_stride in both directions (the overlap carries the blending ramp)")
                size, stride = (tile_size[0] * f, tile_size[1] * f), (tile_stride[0] * f, tile_stride[1] * f)      # pixels, VAE:1224-1226
                weight = torch.zeros(values.shape[1:], dtype=torch.float32, device=self.device)
                border = (tile_size[0] - tile_stride[0], tile_size[1] - tile_stride[1])                            # latent positions
                for h0, h1, w0, w1 in tile_tasks(H, W, size, stride):
                    win = v[:, :, h0:h1, w0:w1].contiguous()
                    self._encode_window(win, values, weight, h0 // f, w0 // f, (h0 == 0, h1 >= H, w0 == 0, w1 >= W), border)
                ops.vae_blend_divide(values, weight)
                self.kernel_launches += 1
            else:
                self._encode_window(v, values, None, 0, 0, (True, True, True, True), (1, 1))
            outs.append(values.to(dtype))
        return torch.stack(outs)


def install(pipe) -> VAE38Encoder:
    """Route ``pipe.vae.encode`` (first-frame conditioning PIPE:490-497, training videos PIPE:374-379) through a VAE38Encoder
    holding the weights of the loaded reference VAE (``vae.install`` does the same for ``decode``)."""
    ref = getattr(pipe, "vae", None)
    if ref is None or not hasattr(ref, "model") or getattr(ref, "z_dim", None) != 48:
        raise ValueError("pipe.vae must be a loaded WanVideoVAE38")
    enc = VAE38Encoder(VAE38, getattr(pipe, "device", "cuda"), enc_dim=ref.model.dim)
    enc.load_state_dict(ref.state_dict())
    ref.encode = enc.encode
    return enc
