"""CPU/PyTorch ORACLE for the umT5 text encoder of Wan2.2.  TEST INFRASTRUCTURE ONLY.

A functional restatement (plain torch ops over a flat ``{name: tensor}`` weight dict) of what the reference's
``WanTextEncoder`` computes and of how the pipeline calls it.  Not part of the product: only ``tests/`` and the bench
tools' CPU-baseline legs may import it; ``fairygen_b200.text_encoder`` never does and has no CPU fallback.

Parity status: PINNED.  ``oracle/make_golden_umt5.py`` runs the real ``WanTextEncoder`` (imported from
``/root/reference/animation`` in the build container) on seeded weights / ids / masks and stores its outputs plus the
relative-position bucket table in ``tests/golden/umt5.npz``; ``tests/test_umt5_oracle.py`` checks this file against them.
The tokenizer (``HuggingfaceTokenizer`` -> ``transformers.AutoTokenizer`` with the google/umt5-xxl vocabulary, TENC:285-330)
needs files that are not in the reference tree; the encoder's contract starts at (ids, mask).

Reference files restated (relative to /root/reference/animation/diffsynth):
  TENC = models/wan_video_text_encoder.py        PIPE = pipelines/wan_video.py
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, Optional

import torch

Weights = Dict[str, torch.Tensor]


@dataclass(frozen=True)
class UMT5Config:                                   # WanTextEncoder.__init__ defaults, TENC:214-223
    vocab: int = 256384
    dim: int = 4096
    dim_attn: int = 4096
    dim_ffn: int = 10240
    num_heads: int = 64
    num_layers: int = 24
    num_buckets: int = 32
    max_dist: int = 128                             # T5RelativeEmbedding default, TENC:150
    eps: float = 1e-6                               # T5LayerNorm default, TENC:27

    @property
    def head_dim(self) -> int:
        return self.dim_attn // self.num_heads


UMT5_XXL = UMT5Config()
TINY = UMT5Config(vocab=96, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2)


def param_shapes(cfg: UMT5Config) -> Dict[str, tuple]:
    """State-dict keys / shapes of WanTextEncoder with shared_pos=False (TENC:233-243, 44-57, 100-104, 133-139)."""
    s = {"token_embedding.weight": (cfg.vocab, cfg.dim), "norm.weight": (cfg.dim,)}
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        s[p + "norm1.weight"] = (cfg.dim,)
        s[p + "norm2.weight"] = (cfg.dim,)
        for n in "qkv":
            s[p + f"attn.{n}.weight"] = (cfg.dim_attn, cfg.dim)
        s[p + "attn.o.weight"] = (cfg.dim, cfg.dim_attn)
        s[p + "ffn.gate.0.weight"] = (cfg.dim_ffn, cfg.dim)
        s[p + "ffn.fc1.weight"] = (cfg.dim_ffn, cfg.dim)
        s[p + "ffn.fc2.weight"] = (cfg.dim, cfg.dim_ffn)
        s[p + "pos_embedding.embedding.weight"] = (cfg.num_buckets, cfg.num_heads)
    return s


def make_weights(cfg: UMT5Config, seed: int = 0) -> Weights:
    """Seeded fp32 weights, tensor by tensor (independent of dict order).  Scales keep the un-scaled T5 scores O(1)."""
    out = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        if name.endswith("norm.weight") or "norm1" in name or "norm2" in name:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name == "token_embedding.weight":
            t = torch.randn(shape, generator=g)
        elif "pos_embedding" in name:
            t = 0.5 * torch.randn(shape, generator=g)
        elif ".attn.q." in name or ".attn.k." in name:
            t = torch.randn(shape, generator=g) * (shape[1] ** -0.5) * (cfg.head_dim ** -0.25)
        else:
            t = torch.randn(shape, generator=g) * (shape[1] ** -0.5)
        out[name] = t
    return out


def make_ids(cfg: UMT5Config, batch: int, seq_len: int, live, seed: int = 3):
    """ids [B, L] (0 = <pad> after the live prefix, 1 = </s> closes it) and the tokenizer's attention mask [B, L]."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(2, cfg.vocab, (batch, seq_len), generator=g)
    mask = torch.zeros(batch, seq_len, dtype=torch.long)
    for b, n in enumerate(live):
        ids[b, n - 1] = 1
        ids[b, n:] = 0
        mask[b, :n] = 1
    return ids, mask


# --------------------------------------------------------------------------------------------
# pieces
# --------------------------------------------------------------------------------------------
def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """GELU, TENC:18-22."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def t5_layer_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """T5LayerNorm.forward, TENC:33-38: no mean subtraction, no bias."""
    x = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    return weight * x.to(weight.dtype)


def relative_position_bucket(rel_pos: torch.Tensor, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """T5RelativeEmbedding._relative_position_bucket with bidirectional=True, TENC:171-193.  rel_pos = key - query."""
    nb = num_buckets // 2
    buckets = (rel_pos > 0).long() * nb
    rel = torch.abs(rel_pos)
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(rel < max_exact, rel, large)


def position_bias(emb: torch.Tensor, lq: int, lk: int, num_buckets: int, max_dist: int) -> torch.Tensor:
    """T5RelativeEmbedding.forward, TENC:159-169 -> [1, heads, lq, lk]."""
    rel = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
    return emb[relative_position_bucket(rel, num_buckets, max_dist)].permute(2, 0, 1).unsqueeze(0).contiguous()


def t5_attention(w: Weights, p: str, cfg: UMT5Config, x: torch.Tensor, mask: Optional[torch.Tensor], pos_bias: torch.Tensor) -> torch.Tensor:
    """T5Attention.forward (self-attention), TENC:59-95: no 1/sqrt(d) scaling, additive bias, masked keys get finfo.min."""
    b, n, c = x.size(0), cfg.num_heads, cfg.head_dim
    q = (x @ w[p + "q.weight"].T).view(b, -1, n, c)
    k = (x @ w[p + "k.weight"].T).view(b, -1, n, c)
    v = (x @ w[p + "v.weight"].T).view(b, -1, n, c)
    bias = x.new_zeros(b, n, q.size(1), k.size(1)) + pos_bias
    if mask is not None:
        bias = bias.masked_fill(mask.view(b, 1, 1, -1) == 0, torch.finfo(x.dtype).min)
    attn = torch.einsum("binc,bjnc->bnij", q, k) + bias
    attn = torch.softmax(attn.float(), dim=-1).type_as(attn)
    out = torch.einsum("bnij,bjnc->binc", attn, v).reshape(b, -1, n * c)
    return out @ w[p + "o.weight"].T


def t5_ffn(w: Weights, p: str, x: torch.Tensor) -> torch.Tensor:
    """T5FeedForward.forward, TENC:108-113: fc2(fc1(x) * gelu(gate(x)))."""
    return ((x @ w[p + "fc1.weight"].T) * gelu_tanh(x @ w[p + "gate.0.weight"].T)) @ w[p + "fc2.weight"].T


def encoder_forward(w: Weights, cfg: UMT5Config, ids: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """WanTextEncoder.forward (eval: dropout = identity), TENC:245-254, blocks per TENC:141-146 with a per-layer bias."""
    x = w["token_embedding.weight"][ids]
    L = x.size(1)
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        e = position_bias(w[p + "pos_embedding.embedding.weight"], L, L, cfg.num_buckets, cfg.max_dist).to(x.dtype)
        x = x + t5_attention(w, p + "attn.", cfg, t5_layer_norm(x, w[p + "norm1.weight"], cfg.eps), mask, e)
        x = x + t5_ffn(w, p + "ffn.", t5_layer_norm(x, w[p + "norm2.weight"], cfg.eps))
    return t5_layer_norm(x, w["norm.weight"], cfg.eps)


def encode_prompt(w: Weights, cfg: UMT5Config, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """WanVideoUnit_PromptEmbedder.encode_prompt after tokenisation, PIPE:404-412: rows from each sample's length on are
    zeroed — in EVERY sample of the batch, as the reference's loop does (``prompt_emb[:, v:] = 0`` for every v)."""
    emb = encoder_forward(w, cfg, ids, mask).clone()
    for v in mask.gt(0).sum(dim=1).long():
        emb[:, int(v):] = 0
    return emb
