"""umT5 text encoder of Wan2.2 on the sm_100a kernels.

Drop-in for ``pipe.text_encoder`` — the reference's ``WanTextEncoder`` (models/wan_video_text_encoder.py:212-257, TENC),
called as ``pipe.text_encoder(ids, mask)`` from ``WanVideoUnit_PromptEmbedder.encode_prompt`` (pipelines/wan_video.py:404-412,
PIPE) once per prompt (positive, negative) before the denoising loop and once per shot in ``batch_inference.py:45-52``.
Same state-dict keys, same ``forward(ids, mask=None) -> [B, L, dim]`` contract; ``encode_prompt`` adds the pipeline's zeroing
of the padded rows.  The tokenizer (``HuggingfaceTokenizer``, TENC:285-330) stays the reference's: it is host string work on
vocabulary files that are not part of the reference tree.

Per block: T5LayerNorm -> ONE GEMM for q|k|v -> ``fgb_t5_attention`` (head_dim 64, additive relative-position bias read from
a per-layer [heads, 2L-1] table instead of a dense [heads, L, L] tensor, masked keys skipped) -> o-projection with the
residual add in the GEMM epilogue -> T5LayerNorm -> ONE GEMM for gate|fc1 -> ``fgb_geglu`` -> fc2 with the residual add in
its epilogue.  Both prompts of a call can go through in one batch (weights are read once).  No CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops
from .ops import BF16, EPI_BIAS, EPI_RESIDUAL


@dataclass(frozen=True)
class UMT5Config:
    """WanTextEncoder.__init__ defaults (TENC:214-223) = google/umt5-xxl encoder."""
    vocab: int = 256384
    dim: int = 4096
    dim_attn: int = 4096
    dim_ffn: int = 10240
    num_heads: int = 64
    num_layers: int = 24
    num_buckets: int = 32
    max_dist: int = 128
    eps: float = 1e-6

    def __post_init__(self):
        if self.dim_attn != self.num_heads * 64:
            raise ValueError(f"fgb_t5_attention is built for head_dim 64: dim_attn {self.dim_attn} != {self.num_heads} * 64")
        if self.dim % 8 or self.dim_ffn % 8:
            raise ValueError("dim and dim_ffn must be multiples of 8 (16-byte rows)")


UMT5_XXL = UMT5Config()


def param_shapes(cfg: UMT5Config) -> Dict[str, tuple]:
    """State-dict keys / shapes of the reference module with shared_pos=False (TENC:233-243)."""
    s = {"token_embedding.weight": (cfg.vocab, cfg.dim), "norm.weight": (cfg.dim,)}
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        s.update({p + "norm1.weight": (cfg.dim,), p + "norm2.weight": (cfg.dim,),
                  p + "attn.q.weight": (cfg.dim_attn, cfg.dim), p + "attn.k.weight": (cfg.dim_attn, cfg.dim),
                  p + "attn.v.weight": (cfg.dim_attn, cfg.dim), p + "attn.o.weight": (cfg.dim, cfg.dim_attn),
                  p + "ffn.gate.0.weight": (cfg.dim_ffn, cfg.dim), p + "ffn.fc1.weight": (cfg.dim_ffn, cfg.dim),
                  p + "ffn.fc2.weight": (cfg.dim, cfg.dim_ffn), p + "pos_embedding.embedding.weight": (cfg.num_buckets, cfg.num_heads)})
    return s


def relative_position_buckets(lq: int, lk: int, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """Bucket of every relative position key - query in [-(lq-1), lk-1] (int32, CPU): the bidirectional T5 bucketing of
    T5RelativeEmbedding._relative_position_bucket (TENC:171-193) — half the buckets per sign; exact below 8, log-spaced up to
    max_dist, saturating beyond.  float32 log on the host, as the reference module evaluates it."""
    rel = torch.arange(-(lq - 1), lk)
    half = num_buckets // 2
    exact = half // 2
    dist = rel.abs()
    log_part = exact + (torch.log(dist.float() / exact) / math.log(max_dist / exact) * (half - exact)).long()
    far = torch.clamp(log_part, max=half - 1)
    return ((rel > 0).long() * half + torch.where(dist < exact, dist, far)).to(torch.int32)


class _Block:
    __slots__ = ("norm1", "wqkv", "wo", "norm2", "wgf", "w2", "pos")


class _Captured:
    """One captured forward: the CUDA graph plus every buffer its kernels were recorded with."""
    __slots__ = ("graph", "ids", "mask", "out", "ws", "bias")


class UMT5Encoder:
    def __init__(self, cfg: UMT5Config = UMT5_XXL, device="cuda", use_graph: bool = True):
        """use_graph: the ~200 launches of a forward are a few microseconds of GPU work each at prompt sizes, so the forward
        is captured once per (batch, length, masked?) into a CUDA graph and replayed (the launch-bound inner loop)."""
        self.cfg = cfg
        self.use_graph = use_graph
        self._graphs: Dict[tuple, _Captured] = {}
        self.device = torch.device(device)
        ops.context(self.device)       # raises off-GPU: there is no CPU path
        self.loaded = False
        self.blocks = []
        self.kernel_launches = 0
        self._bias_cache: Dict[int, torch.Tensor] = {}
        self._ws: Dict[int, Dict[str, torch.Tensor]] = {}

    # ------------------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        """Keys of the reference module (models_t5_umt5-xxl-enc-bf16.pth).  q|k|v and gate|fc1 are stored row-concatenated."""
        cfg, dev = self.cfg, self.device
        want = param_shapes(cfg)
        missing = [k for k in want if k not in sd]
        extra = [k for k in sd if k not in want]
        if missing or (strict and extra):
            raise KeyError(f"umT5 state dict mismatch: missing {missing[:4]}{'...' if len(missing) > 4 else ''}, "
                           f"unexpected {extra[:4]}{'...' if len(extra) > 4 else ''}")
        for k, shape in want.items():
            if tuple(sd[k].shape) != shape:
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {shape}")
        g = lambda k: sd[k].detach().to(device=dev, dtype=BF16).contiguous()  # noqa: E731
        self.token_embedding = g("token_embedding.weight")
        self.norm = g("norm.weight")
        self.blocks = []
        for i in range(cfg.num_layers):
            p = f"blocks.{i}."
            b = _Block()
            b.norm1, b.norm2 = g(p + "norm1.weight"), g(p + "norm2.weight")
            b.wqkv = torch.cat([g(p + f"attn.{n}.weight") for n in "qkv"], dim=0)
            b.wo = g(p + "attn.o.weight")
            b.wgf = torch.cat([g(p + "ffn.gate.0.weight"), g(p + "ffn.fc1.weight")], dim=0)
            b.w2 = g(p + "ffn.fc2.weight")
            b.pos = g(p + "pos_embedding.embedding.weight")
            self.blocks.append(b)
        self._bias_cache.clear()
        self._graphs.clear()           # captured graphs hold pointers to the previous weights
        self.loaded = True

    def _bias_tables(self, L: int) -> torch.Tensor:
        """[layers, heads, 2L-1] fp32: every layer's position bias by relative position (built once per sequence length)."""
        tab = self._bias_cache.get(L)
        if tab is None:
            cfg = self.cfg
            buckets = relative_position_buckets(L, L, cfg.num_buckets, cfg.max_dist).to(self.device)
            tab = torch.empty(cfg.num_layers, cfg.num_heads, 2 * L - 1, dtype=torch.float32, device=self.device)
            for i, b in enumerate(self.blocks):
                ops.t5_bias_table(b.pos, buckets, tab[i])
            self.kernel_launches += cfg.num_layers
            if len(self._bias_cache) >= 8:        # prompts come in many lengths (encode_prompt trims to the live tokens)
                self._bias_cache.pop(next(iter(self._bias_cache)))
            self._bias_cache[L] = tab
        return tab

    def _workspace(self, rows: int) -> Dict[str, torch.Tensor]:
        ws = self._ws.get(rows)
        if ws is None:
            cfg = self.cfg
            e = lambda *s: torch.empty(*s, dtype=BF16, device=self.device)  # noqa: E731
            ws = dict(x=e(rows, cfg.dim), n=e(rows, cfg.dim), qkv=e(rows, 3 * cfg.dim_attn), o=e(rows, cfg.dim_attn),
                      gf=e(rows, 2 * cfg.dim_ffn), h=e(rows, cfg.dim_ffn))
            self._ws = {rows: ws}
        return ws

    # ------------------------------------------------------------------------------------------------------------
    def _launch(self, ids_dev, key_mask, B: int, ws, bias, out) -> None:
        """The kernels of one forward, in order, on the current stream; allocates nothing (capturable)."""
        cfg = self.cfg
        x, n, qkv, o, gf, h = (ws[k] for k in ("x", "n", "qkv", "o", "gf", "h"))
        da, H = cfg.dim_attn, cfg.num_heads
        ops.embedding_rows(self.token_embedding, ids_dev, x)
        for i, b in enumerate(self.blocks):
            ops.t5_layer_norm(x, n, cfg.eps, b.norm1)
            ops.gemm(n, b.wqkv, None, qkv, EPI_BIAS)
            ops.t5_attention(qkv[:, :da], qkv[:, da:2 * da], qkv[:, 2 * da:], o, B, H, bias=bias[i], key_mask=key_mask)
            ops.gemm(o, b.wo, None, x, EPI_RESIDUAL)                     # x += o(attn)            TENC:144
            ops.t5_layer_norm(x, n, cfg.eps, b.norm2)
            ops.gemm(n, b.wgf, None, gf, EPI_BIAS)
            ops.geglu(gf, h)
            ops.gemm(h, b.w2, None, x, EPI_RESIDUAL)                     # x += fc2(fc1 * gelu(gate))  TENC:145
        ops.t5_layer_norm(x, out, cfg.eps, self.norm)
        self.kernel_launches += 2 + 8 * cfg.num_layers

    def _capture(self, key, ids_dev, key_mask, B: int, L: int) -> _Captured:
        cfg, dev = self.cfg, self.device
        c = _Captured()
        c.ids = ids_dev.clone()
        c.mask = None if key_mask is None else key_mask.clone()
        c.out = torch.empty(B * L, cfg.dim, dtype=BF16, device=dev)
        e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)  # noqa: E731
        rows = B * L
        c.ws = dict(x=e(rows, cfg.dim), n=e(rows, cfg.dim), qkv=e(rows, 3 * cfg.dim_attn), o=e(rows, cfg.dim_attn),
                    gf=e(rows, 2 * cfg.dim_ffn), h=e(rows, cfg.dim_ffn))
        c.bias = self._bias_tables(L)
        self._launch(c.ids, c.mask, B, c.ws, c.bias, c.out)        # eager once: one-time kernel attributes are set outside capture
        torch.cuda.synchronize(dev)
        c.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(c.graph):
            self._launch(c.ids, c.mask, B, c.ws, c.bias, c.out)
        if len(self._graphs) >= 16:
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = c
        return c

    @torch.no_grad()
    def forward(self, ids: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """WanTextEncoder.forward (TENC:245-254): ids [B, L] int, mask [B, L] (0 = padded key) or None -> [B, L, dim] bf16."""
        if not self.loaded:
            raise RuntimeError("UMT5Encoder.forward called before load_state_dict")
        cfg, dev = self.cfg, self.device
        if ids.dim() != 2 or ids.is_floating_point():
            raise ValueError(f"ids must be an integer tensor [batch, seq_len], got {ids.dtype} {tuple(ids.shape)}")
        B, L = ids.shape
        if int(ids.min()) < 0 or int(ids.max()) >= cfg.vocab:      # nn.Embedding raises IndexError for these
            raise IndexError(f"token id out of range [0, {cfg.vocab})")
        key_mask = None
        if mask is not None:
            if tuple(mask.shape) != (B, L):
                raise ValueError(f"mask must be [batch, seq_len] = [{B}, {L}], got {tuple(mask.shape)}")
            if bool((mask != 0).sum(dim=1).eq(0).any()):
                raise ValueError("mask leaves a sample without any key (the tokenizer always emits </s>)")
            key_mask = (mask != 0).to(device=dev, dtype=torch.uint8).contiguous()
        ids_dev = ids.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        if self.use_graph and not torch.cuda.is_current_stream_capturing():
            key = (B, L, key_mask is not None)
            c = self._graphs.get(key)
            if c is None:
                c = self._capture(key, ids_dev, key_mask, B, L)
            c.ids.copy_(ids_dev)
            if key_mask is not None:
                c.mask.copy_(key_mask)
            c.graph.replay()
            self.kernel_launches += 2 + 8 * cfg.num_layers
            return c.out.clone().view(B, L, cfg.dim)      # the graph's output buffer is overwritten by the next call
        out = torch.empty(B * L, cfg.dim, dtype=BF16, device=dev)
        self._launch(ids_dev, key_mask, B, self._workspace(B * L), self._bias_tables(L), out)
        return out.view(B, L, cfg.dim)

    __call__ = forward

    def _encode_live(self, ids: torch.Tensor, mask: torch.Tensor):
        """Encoder output for the first max(live) positions only, placed in a zeroed [B, L, dim] tensor.  Rows of live tokens
        never see the padded positions (masked as keys), so trimming the padded tail changes none of them."""
        if ids.dim() != 2 or tuple(mask.shape) != tuple(ids.shape):
            raise ValueError(f"ids and mask must both be [batch, seq_len], got {tuple(ids.shape)} and {tuple(mask.shape)}")
        lens = (mask != 0).sum(dim=1).tolist()
        B, L = ids.shape
        n = max(int(v) for v in lens)
        if n == 0 or bool((mask[:, :n] != 0).sum(dim=1).ne(torch.tensor(lens, device=mask.device)).any()):
            n = L          # not the tokenizer's prefix mask: no trimming
        n = min(L, -(-n // 32) * 32)   # a few extra masked positions; prompt lengths then share workspaces / captured graphs
        emb = torch.zeros(B, L, self.cfg.dim, dtype=BF16, device=self.device)
        emb[:, :n] = self.forward(ids[:, :n], mask[:, :n])
        return emb, lens

    @torch.no_grad()
    def encode_prompt(self, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """encode_prompt after tokenisation (PIPE:404-412): the encoder output with the rows from each sample's length on set
        to zero — in every sample of the batch, as the reference's ``prompt_emb[:, v:] = 0`` loop does (the reference calls
        it with one prompt at a time).  Only the live prefix is computed."""
        emb, lens = self._encode_live(ids, mask)
        for v in lens:
            emb[:, int(v):] = 0
        return emb

    @torch.no_grad()
    def encode_prompts(self, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """Several prompts (e.g. positive + negative of one call) as ONE batch, each treated as its own reference
        ``encode_prompt`` call: sample b is zeroed from ITS length on.  The weights are read once for all of them."""
        emb, lens = self._encode_live(ids, mask)
        for b, v in enumerate(lens):
            emb[b, int(v):] = 0
        return emb


def config_of(module) -> UMT5Config:
    """UMT5Config of a reference ``WanTextEncoder`` instance (its constructor arguments, TENC:225-243)."""
    return UMT5Config(vocab=module.token_embedding.num_embeddings, dim=module.dim, dim_attn=module.dim_attn, dim_ffn=module.dim_ffn,
                      num_heads=module.num_heads, num_layers=module.num_layers, num_buckets=module.num_buckets)


def install(pipe) -> UMT5Encoder:
    """Replace ``pipe.text_encoder`` (a loaded reference WanTextEncoder, PIPE:133) by a UMT5Encoder holding the same weights.
    ``pipe.text_encoder(ids, mask)`` (PIPE:409) then runs on the kernels; the reference module is dropped from the pipeline's
    children, so its offload / onload bookkeeping (base_pipeline.py:146-168) no longer touches the encoder."""
    ref = getattr(pipe, "text_encoder", None)
    if ref is None or getattr(ref, "shared_pos", False):
        raise ValueError("pipe.text_encoder must be a loaded WanTextEncoder with per-layer position embeddings (shared_pos=False)")
    enc = UMT5Encoder(config_of(ref), getattr(pipe, "device", "cuda"))
    enc.load_state_dict(ref.state_dict())
    if isinstance(pipe, torch.nn.Module):
        delattr(pipe, "text_encoder")               # unregister the child module before storing a plain object
        object.__setattr__(pipe, "text_encoder", enc)
    else:
        pipe.text_encoder = enc
    return enc
