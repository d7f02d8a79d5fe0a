"""Generate tests/golden/umt5.npz by running the REAL reference text encoder (read-only, /root/reference/animation).

Runs only in the build container.  Weights, ids and masks are regenerated from seeds by ``umt5_oracle.make_weights /
make_ids``; only the reference's OUTPUTS are stored (fp32, CPU):

    python oracle/make_golden_umt5.py            # rewrites tests/golden/umt5.npz
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

from diffsynth.models import wan_video_text_encoder as te  # noqa: E402

from oracle import umt5_oracle as u  # noqa: E402

CASES = {"short": (1, 40, (13,)), "pair": (2, 48, (48, 7)), "long": (1, 200, (170,))}   # (batch, seq_len, live tokens)


@torch.no_grad()
def main():
    cfg = u.TINY
    w = u.make_weights(cfg, seed=0)
    enc = te.WanTextEncoder(vocab=cfg.vocab, dim=cfg.dim, dim_attn=cfg.dim_attn, dim_ffn=cfg.dim_ffn, num_heads=cfg.num_heads,
                            num_layers=cfg.num_layers, num_buckets=cfg.num_buckets, shared_pos=False).eval()
    enc.load_state_dict(w, strict=True)
    out = {}
    for name, (b, L, live) in CASES.items():
        ids, mask = u.make_ids(cfg, b, L, live, seed=3)
        y = enc(ids, mask)
        out[name] = y.float().numpy()
        # the pipeline's call (PIPE:404-412), restated on the reference module's output
        z = y.clone()
        for v in mask.gt(0).sum(dim=1).long():
            z[:, v:] = 0
        out[name + "_prompt"] = z.float().numpy()
    out["nomask"] = enc(u.make_ids(cfg, 1, 24, (24,), seed=5)[0]).float().numpy()
    rel = torch.arange(-600, 601)
    out["buckets"] = te.T5RelativeEmbedding(cfg.num_buckets, cfg.num_heads, bidirectional=True)._relative_position_bucket(rel).numpy()
    blk = enc.blocks[1]
    out["pos_bias_l1"] = blk.pos_embedding(20, 20).float().numpy()
    x = torch.randn(3, 5, cfg.dim, generator=torch.Generator().manual_seed(8))
    out["layer_norm"] = blk.norm1(x).numpy()
    out["ffn"] = blk.ffn(x).numpy()
    path = os.path.join(REPO, "tests", "golden", "umt5.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
