"""Model dimensions of the DiT on the hot path (reference: animation/diffsynth/configs/model_configs.py:290-295)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class WanDiTConfig:
    dim: int = 3072
    in_dim: int = 48
    ffn_dim: int = 14336
    out_dim: int = 48
    text_dim: int = 4096
    freq_dim: int = 256
    eps: float = 1e-6
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    num_heads: int = 24
    num_layers: int = 30
    seperated_timestep: bool = True  # (sic) attribute name used by the reference WanModel

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads

    def validate(self) -> None:
        if self.head_dim != 128:
            raise ValueError(f"fairygen_b200 kernels are specialised for head_dim 128, got {self.head_dim}")
        if tuple(self.patch_size) != (1, 2, 2):
            raise ValueError(f"patch_size {self.patch_size} unsupported; the hot path is patch (1,2,2)")
        if self.dim % 256 != 0:
            raise ValueError("dim must be a multiple of 256")
        if self.ffn_dim % 8 or self.text_dim % 8 or self.freq_dim % 8:
            raise ValueError("ffn_dim, text_dim and freq_dim must be multiples of 8")

    @staticmethod
    def from_state_dict(sd) -> "WanDiTConfig":
        """Infer the dimensions from reference WanModel state-dict shapes (DIT:271-336)."""
        dim, in_dim, pt, ph, pw = sd["patch_embedding.weight"].shape
        layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
        out_dim = sd["head.head.weight"].shape[0] // (pt * ph * pw)
        return WanDiTConfig(
            dim=dim, in_dim=in_dim, ffn_dim=sd["blocks.0.ffn.0.weight"].shape[0], out_dim=out_dim,
            text_dim=sd["text_embedding.0.weight"].shape[1], freq_dim=sd["time_embedding.0.weight"].shape[1],
            patch_size=(pt, ph, pw), num_heads=dim // 128, num_layers=layers,
        )

    @staticmethod
    def from_module(dit) -> "WanDiTConfig":
        """Read the dimensions off a reference ``WanModel`` instance (the weight container we sit behind)."""
        blk = dit.blocks[0]
        return WanDiTConfig(
            dim=dit.dim, in_dim=dit.in_dim, ffn_dim=blk.ffn_dim, out_dim=dit.head.head.out_features // 4,
            text_dim=dit.text_embedding[0].in_features, freq_dim=dit.freq_dim, eps=blk.norm1.eps,
            patch_size=tuple(dit.patch_size), num_heads=blk.num_heads, num_layers=len(dit.blocks),
            seperated_timestep=bool(dit.seperated_timestep),
        )


TI2V_5B = WanDiTConfig()


def counted_flops(cfg: WanDiTConfig, tokens: int, text_len: int = 512) -> float:
    """Algorithmic FLOPs of one DiT forward; step-invariant work is excluded."""
    d, f, s = cfg.dim, cfg.ffn_dim, tokens
    per_block = 12 * s * d * d + 4 * s * d * f + 4 * s * s * d + 4 * s * text_len * d
    io = cfg.out_dim * cfg.patch_size[0] * cfg.patch_size[1] * cfg.patch_size[2]
    return cfg.num_layers * per_block + 2 * (2 * s * d * io)
