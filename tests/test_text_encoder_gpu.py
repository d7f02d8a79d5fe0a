"""umT5 text encoder on the B200 against the pinned oracle (oracle/umt5_oracle.py, itself checked against the real reference
module in tests/test_umt5_oracle.py): every kernel alone, the tiny encoder on the golden cases (also directly against the
reference's stored outputs), one layer at the umt5-xxl dimensions, and the error paths."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
GOLD = os.path.join(os.path.dirname(__file__), "golden", "umt5.npz")
CASES = {"short": (1, 40, (13,)), "pair": (2, 48, (48, 7)), "long": (1, 200, (170,))}


@pytest.fixture(scope="module")
def env():
    from fairygen_b200 import ops, text_encoder
    from oracle import umt5_oracle as u
    torch.cuda.set_device(0)
    ops.context(torch.device("cuda", 0))
    return ops, text_encoder, u


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


def test_embedding_layer_norm_geglu(env):
    ops, _, u = env
    table = rnd(300, 256, seed=1)
    ids = torch.tensor([0, 299, 7, 7, 123], device="cuda")
    out = torch.empty(5, 256, dtype=BF, device="cuda")
    ops.embedding_rows(table, ids, out)
    assert torch.equal(out, table[ids])
    for rows, dim in [(5, 128), (513, 4096), (40, 1000)]:
        x, w = rnd(rows, dim, seed=2, scale=3.0), 1 + rnd(dim, seed=3, scale=0.2)
        y = torch.empty_like(x)
        ops.t5_layer_norm(x, y, 1e-6, w)
        want = u.t5_layer_norm(x, w, 1e-6)          # the reference's own bf16 rounding points
        assert rel_l2(y, want) < 2e-3
    gf = rnd(37, 2 * 264, seed=4, scale=2.0)
    h = torch.empty(37, 264, dtype=BF, device="cuda")
    ops.geglu(gf, h)
    ops.sync_check()
    assert rel_l2(h, gf[:, 264:].float() * u.gelu_tanh(gf[:, :264].float())) < 4e-3


@pytest.mark.parametrize("batch,s,heads,live", [(1, 40, 2, (13,)), (2, 200, 3, (200, 129)), (1, 512, 4, (300,)), (2, 33, 1, (1, 33))])
def test_t5_attention_kernel(env, batch, s, heads, live):
    ops, te, u = env
    q, k, v = (rnd(batch * s, heads * 64, seed=i, scale=sc) for i, sc in ((1, 0.4), (2, 0.4), (3, 1.0)))
    emb = rnd(32, heads, seed=4, scale=0.7)
    mask = torch.zeros(batch, s, dtype=torch.uint8, device="cuda")
    for b, n in enumerate(live):
        mask[b, :n] = 1
    tab = torch.empty(heads, 2 * s - 1, dtype=torch.float32, device="cuda")
    ops.t5_bias_table(emb, te.relative_position_buckets(s, s).cuda(), tab)
    dense = u.position_bias(emb.float().cpu(), s, s, 32, 128).cuda()          # [1, heads, s, s]
    idx = (torch.arange(s)[None, :] - torch.arange(s)[:, None] + s - 1).cuda()
    assert torch.equal(tab[:, idx], dense[0])
    out = torch.full((batch * s, heads * 64), float("nan"), dtype=BF, device="cuda")
    ops.t5_attention(q, k, v, out, batch, heads, bias=tab, key_mask=mask)
    ops.sync_check()
    qf, kf, vf = (t.float().view(batch, s, heads, 64) for t in (q, k, v))
    sc = torch.einsum("binc,bjnc->bnij", qf, kf) + dense
    sc = sc.masked_fill(mask.view(batch, 1, 1, s) == 0, float("-inf"))
    want = torch.einsum("bnij,bjnc->binc", torch.softmax(sc, dim=-1), vf).reshape(batch * s, heads * 64)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, want) < 4e-3
    # no bias, no mask
    ops.t5_attention(q, k, v, out, batch, heads)
    ops.sync_check()
    want = torch.einsum("bnij,bjnc->binc", torch.softmax(torch.einsum("binc,bjnc->bnij", qf, kf), dim=-1), vf).reshape(batch * s, heads * 64)
    assert rel_l2(out, want) < 4e-3


def _tiny(env):
    ops, te, u = env
    ocfg = u.TINY
    cfg = te.UMT5Config(vocab=ocfg.vocab, dim=ocfg.dim, dim_attn=ocfg.dim_attn, dim_ffn=ocfg.dim_ffn, num_heads=ocfg.num_heads,
                        num_layers=ocfg.num_layers)
    w = u.make_weights(ocfg, seed=0)
    enc = te.UMT5Encoder(cfg, "cuda")
    enc.load_state_dict(w)
    return ocfg, w, enc


@pytest.mark.parametrize("name", list(CASES))
def test_tiny_encoder_vs_oracle_and_reference_golden(env, name):
    ops, te, u = env
    ocfg, w, enc = _tiny(env)
    b, L, live = CASES[name]
    ids, mask = u.make_ids(ocfg, b, L, live, seed=3)
    out = enc(ids, mask)
    ops.sync_check()
    w16 = {k: v.to(BF).float().cuda() for k, v in w.items()}          # the oracle on the bf16-rounded weights the GPU holds
    want = u.encoder_forward(w16, ocfg, ids.cuda(), mask.cuda())
    gold = torch.from_numpy(np.load(GOLD)[name]).cuda()               # the reference module itself (fp32 weights)
    # rows of live tokens are what the pipeline keeps; padded-query rows are compared too (the reference computes them)
    assert rel_l2(out, want) < 1e-2, rel_l2(out, want)
    assert rel_l2(out, gold) < 1.5e-2, rel_l2(out, gold)
    emb = enc.encode_prompt(ids, mask)
    ops.sync_check()
    gold_p = torch.from_numpy(np.load(GOLD)[name + "_prompt"]).cuda()
    assert not emb[:, min(live):].float().abs().max() > 0
    assert rel_l2(emb, gold_p) < 1.5e-2
    print(f"umT5 tiny {name}: vs oracle {rel_l2(out, want):.3e}, vs reference golden {rel_l2(out, gold):.3e}")


def test_encode_prompts_equals_one_reference_call_per_prompt(env):
    """Positive + negative prompt as one batch == two encode_prompt calls (each zeroed from its own length on)."""
    ops, te, u = env
    ocfg, w, enc = _tiny(env)
    ids, mask = u.make_ids(ocfg, 2, 64, (50, 9), seed=6)
    both = enc.encode_prompts(ids, mask)
    ops.sync_check()
    w16 = {k: v.to(BF).float().cuda() for k, v in w.items()}
    for b in range(2):
        want = u.encode_prompt(w16, ocfg, ids[b:b + 1].cuda(), mask[b:b + 1].cuda())
        alone = enc.encode_prompt(ids[b:b + 1], mask[b:b + 1])
        assert rel_l2(both[b:b + 1], want) < 1e-2
        assert rel_l2(both[b:b + 1], alone) < 2e-3       # batch of 2 vs batch of 1: same kernels, other GEMM tile fill
        assert not both[b, int(mask[b].sum()):].float().abs().max() > 0
    # a mask that is not a prefix (never produced by the tokenizer) is computed untrimmed
    odd = mask.clone()
    odd[0, 60] = 1
    full = enc.encode_prompts(ids, odd)
    want = u.encoder_forward(w16, ocfg, ids.cuda(), odd.cuda())
    assert rel_l2(full[1, :9], want[1, :9]) < 1e-2 and rel_l2(full[0, :51], want[0, :51]) < 1e-2


def test_encoder_without_mask_and_cached_bias(env):
    ops, te, u = env
    ocfg, w, enc = _tiny(env)
    ids, _ = u.make_ids(ocfg, 1, 24, (24,), seed=5)
    out = enc(ids)
    out2 = enc(ids.cuda())                                            # device ids, bias table from the cache
    ops.sync_check()
    gold = torch.from_numpy(np.load(GOLD)["nomask"]).cuda()
    assert rel_l2(out, gold) < 1.5e-2 and torch.equal(out, out2)


def test_one_layer_at_umt5_xxl_dimensions(env):
    """dim 4096, 64 heads x 64, ffn 10240, 512 tokens x 2 prompts: the production tile shapes of every kernel, 1 layer."""
    ops, te, u = env
    ocfg = u.UMT5Config(vocab=1000, num_layers=1)
    cfg = te.UMT5Config(vocab=1000, num_layers=1)
    w = u.make_weights(ocfg, seed=1)
    enc = te.UMT5Encoder(cfg, "cuda")
    enc.load_state_dict(w)
    ids, mask = u.make_ids(ocfg, 2, 512, (77, 300), seed=4)
    out = enc(ids, mask)
    ops.sync_check()
    w16 = {k: v.to(BF).float().cuda() for k, v in w.items()}
    want = u.encoder_forward(w16, ocfg, ids.cuda(), mask.cuda())
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, want) < 1e-2, rel_l2(out, want)


def test_captured_graph_equals_eager_launches(env):
    """The CUDA-graph replay (default) and the plain launch sequence give the same bits, call after call, for new ids / masks
    of a captured shape and for a new shape."""
    ops, te, u = env
    ocfg, w, enc = _tiny(env)
    eager = te.UMT5Encoder(enc.cfg, "cuda", use_graph=False)
    eager.load_state_dict(w)
    for seed, (b, L, live) in enumerate([(2, 64, (40, 64)), (2, 64, (64, 3)), (1, 40, (13,)), (2, 64, (5, 6))]):
        ids, mask = u.make_ids(ocfg, b, L, live, seed=20 + seed)
        got, want = enc(ids, mask), eager(ids, mask)
        ops.sync_check()
        assert torch.equal(got, want), (seed, rel_l2(got, want))
    assert len(enc._graphs) == 2 and not eager._graphs


def test_install_replaces_the_pipeline_text_encoder(env):
    """text_encoder.install(pipe): the reference module (a stand-in with its attributes and state dict) is swapped for the
    kernel encoder, called exactly as PIPE:409 calls it."""
    ops, te, u = env
    ocfg = u.TINY
    w = u.make_weights(ocfg, seed=0)

    class RefEncoder(torch.nn.Module):          # what install() reads off a WanTextEncoder (TENC:225-243)
        def __init__(self):
            super().__init__()
            self.dim, self.dim_attn, self.dim_ffn = ocfg.dim, ocfg.dim_attn, ocfg.dim_ffn
            self.num_heads, self.num_layers, self.num_buckets, self.shared_pos = ocfg.num_heads, ocfg.num_layers, ocfg.num_buckets, False
            self.token_embedding = torch.nn.Embedding(ocfg.vocab, ocfg.dim)

        def state_dict(self, *a, **k):
            return dict(w)

    class Pipe(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.text_encoder = RefEncoder()
            self.device = "cuda"

    pipe = Pipe()
    enc = te.install(pipe)
    assert pipe.text_encoder is enc and "text_encoder" not in dict(pipe.named_children())
    ids, mask = u.make_ids(ocfg, 1, 40, (13,), seed=3)
    out = pipe.text_encoder(ids.cuda(), mask.cuda())
    ops.sync_check()
    assert rel_l2(out, torch.from_numpy(np.load(GOLD)["short"]).cuda()) < 1.5e-2
    pipe.text_encoder = None
    with pytest.raises(ValueError):
        te.install(pipe)


def test_text_encoder_rejects_bad_arguments(env):
    ops, te, u = env
    ocfg, w, enc = _tiny(env)
    with pytest.raises(RuntimeError):
        te.UMT5Encoder(enc.cfg, "cuda")(torch.zeros(1, 4, dtype=torch.long))          # not loaded
    with pytest.raises(IndexError):
        enc(torch.full((1, 4), ocfg.vocab, dtype=torch.long))
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 4, dtype=torch.long), torch.zeros(1, 4, dtype=torch.long))  # nothing to attend to
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 4, dtype=torch.long), torch.ones(1, 5, dtype=torch.long))
    with pytest.raises(KeyError):
        te.UMT5Encoder(enc.cfg, "cuda").load_state_dict({k: v for k, v in w.items() if k != "norm.weight"})
    with pytest.raises(ValueError):
        te.UMT5Config(dim_attn=4096, num_heads=32)                                      # head_dim 128: not this kernel
    with pytest.raises(ValueError):
        ops.t5_attention(rnd(8, 64), rnd(8, 64), rnd(8, 64), torch.empty(8, 64, dtype=BF, device="cuda"), 1, 1,
                         bias=torch.zeros(1, 14, device="cuda"))
