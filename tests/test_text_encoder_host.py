"""Host logic of the umT5 encoder (fairygen_b200/text_encoder.py) on the CPU: the kernels are replaced by plain-torch statements
of their contracts (include/fairygen_b200.h) and the orchestration — fused q|k|v and gate|fc1 weights, per-layer bias tables by
relative position, residual epilogues, live-prefix trimming, per-prompt zeroing — must reproduce the pinned oracle and the
reference's stored outputs.  The kernels themselves are covered by tests/test_text_encoder_gpu.py."""
import os

import numpy as np
import pytest
import torch

from oracle import umt5_oracle as u

BF = torch.bfloat16
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "umt5.npz"))
CASES = {"short": (1, 40, (13,)), "pair": (2, 48, (48, 7)), "long": (1, 200, (170,))}


def _emulated_ops(monkeypatch):
    from fairygen_b200 import ops

    def embedding_rows(table, ids, out):
        out.copy_(table[ids])
        return out

    def t5_layer_norm(x, out, eps, weight):
        xf = x.float()
        out.copy_((weight.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(BF).float()).to(BF))
        return out

    def gemm(a, w, bias, out, epilogue=ops.EPI_BIAS, *args, **kw):
        y = (a.float() @ w.float().T + (0 if bias is None else bias.float())).to(BF).float()
        out.copy_(((out.float() + y) if epilogue == ops.EPI_RESIDUAL else y).to(BF))
        return out

    def geglu(gate_fc1, out):
        f = gate_fc1.shape[1] // 2
        out.copy_((gate_fc1[:, f:].float() * u.gelu_tanh(gate_fc1[:, :f].float()).to(BF).float()).to(BF))
        return out

    def t5_bias_table(emb, bucket_of_rel, out):
        out.copy_(emb.float()[bucket_of_rel.long()].T)
        return out

    def t5_attention(q, k, v, out, batch, heads, bias=None, key_mask=None):
        s_q, s_kv = q.shape[0] // batch, k.shape[0] // batch
        qf, kf, vf = (t.float().reshape(batch, -1, heads, 64) for t in (q, k, v))
        sc = torch.einsum("binc,bjnc->bnij", qf, kf)
        if bias is not None:
            idx = torch.arange(s_kv)[None, :] - torch.arange(s_q)[:, None] + s_q - 1
            sc = sc + bias[:, idx][None]
        if key_mask is not None:
            sc = sc.masked_fill(key_mask.view(batch, 1, 1, s_kv) == 0, float("-inf"))
        out.copy_(torch.einsum("bnij,bjnc->binc", torch.softmax(sc, -1), vf).reshape(batch * s_q, heads * 64).to(BF))
        return out

    for name, fn in list(locals().items()):
        if callable(fn) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(ops, "context", lambda device: None)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.fixture()
def encoder(monkeypatch):
    _emulated_ops(monkeypatch)
    from fairygen_b200 import text_encoder as te
    ocfg = u.TINY
    cfg = te.UMT5Config(vocab=ocfg.vocab, dim=ocfg.dim, dim_attn=ocfg.dim_attn, dim_ffn=ocfg.dim_ffn, num_heads=ocfg.num_heads,
                        num_layers=ocfg.num_layers)
    w = u.make_weights(ocfg, seed=0)
    enc = te.UMT5Encoder(cfg, "cpu", use_graph=False)          # CUDA graphs need a GPU; the launch sequence is the same code
    enc.load_state_dict(w)
    return enc, {k: v.to(BF).float() for k, v in w.items()}


@pytest.mark.parametrize("name", list(CASES))
def test_forward_and_prompt_embedding(encoder, name):
    enc, w16 = encoder
    b, L, live = CASES[name]
    ids, mask = u.make_ids(u.TINY, b, L, live, seed=3)
    out = enc(ids, mask)
    assert out.shape == (b, L, u.TINY.dim) and out.dtype == BF
    assert rel(out.float(), u.encoder_forward(w16, u.TINY, ids, mask)) < 1e-2
    assert rel(out.float(), torch.from_numpy(GOLD[name])) < 1.5e-2
    emb = enc.encode_prompt(ids, mask)                          # live prefix only, zeroed from the shortest length on
    assert rel(emb.float(), torch.from_numpy(GOLD[name + "_prompt"])) < 1.5e-2
    assert min(live) == L or not emb[:, min(live):].float().abs().max() > 0
    each = enc.encode_prompts(ids, mask)                        # every prompt zeroed from ITS length on
    for i, n in enumerate(live):
        assert rel(each[i, :n].float(), u.encoder_forward(w16, u.TINY, ids[i:i + 1], mask[i:i + 1])[0, :n]) < 1e-2
        assert n == L or not each[i, n:].float().abs().max() > 0


def test_no_mask_and_errors(encoder):
    enc, _ = encoder
    ids, _ = u.make_ids(u.TINY, 1, 24, (24,), seed=5)
    assert rel(enc(ids).float(), torch.from_numpy(GOLD["nomask"])) < 1.5e-2
    with pytest.raises(IndexError):
        enc(torch.full((1, 4), u.TINY.vocab, dtype=torch.long))
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 4, dtype=torch.long), torch.zeros(1, 4, dtype=torch.long))
