"""Per-kernel parity on the B200: every C-ABI entry point against the oracle's torch restatement of the
same reference op (fp32 math on the same bf16 inputs), plus bit-exact checks where the op is pure
rounding arithmetic (scheduler step, layout kernels) against the golden vectors from the real reference."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def env():
    from fairygen_b200 import ops
    from oracle import wan_dit_oracle as o
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    ops.context(torch.device("cuda", 0))  # raises if the CUDA extension is missing / not sm_100
    return ops, o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


def assert_head_bound(got, exact):
    """Per-head maxima left by the norm kernels against fgb_head_norm_max of their output (different summation order only)."""
    assert torch.allclose(got, exact, rtol=1e-5), (got, exact)


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (2, 768, 256), (200, 192, 192), (1000, 3072, 3072), (513, 1536, 256),
                                   (300, 512, 4096), (2049, 14336, 3072),
                                   # the 2-CTA kernel (m >= 256): exact tiles, ragged rows / columns / k, one tile, many waves
                                   (256, 256, 64), (257, 200, 72), (384, 8, 8), (511, 264, 136), (20000, 3072, 512),
                                   # half-width tail: 81 tiles = one wave of 74 + 7 tiles as 14 half-width items; 1284 = 17 waves + 26
                                   (6820, 768, 512), (27280, 3072, 128)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_epilogues(env, m, n, k, epi):
    ops, o = env
    if epi != 0 and m * n * k > 3.2e10:
        pytest.skip("large shape checked with the plain epilogue only")
    a, w, bias = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=1 / math.sqrt(k)), rnd(n, seed=3, scale=0.5)
    g0, g1, c0 = rnd(n, seed=4), rnd(n, seed=5), rnd(m, n, seed=6)
    rows0 = m // 3
    out = c0.clone()
    ops.gemm(a, w, bias, out, epi, g0, g1, rows0)
    ops.sync_check()
    y = (a.float() @ w.float().T + bias.float()).to(BF).float()     # nn.Linear output in bf16
    if epi == 0:
        ref = y
    elif epi == 1:
        ref = F.gelu(y, approximate="tanh")
    elif epi == 2:
        gate = torch.where((torch.arange(m, device="cuda") < rows0)[:, None], g0.float()[None], g1.float()[None])
        ref = c0.float() + (gate * y).to(BF).float()
    else:
        ref = c0.float() + y
    assert rel_l2(out, ref.to(BF)) < 2e-3


@pytest.mark.parametrize("m,n,k", [(6820, 3072, 3072),     # Ulysses SP4 rank: 324 tiles = 4 waves + 28 tail tiles over 74 CTA pairs
                                   (1000, 768, 512),       # 12 tiles < 74 pairs: no whole wave, 2 units per tile, idle pairs
                                   (2561, 9216, 1000),     # ragged m / k; 396 tiles = 5 waves + 26; last unit of a tile is short
                                   (19000, 1024, 1536),    # 300 tiles = 4 waves + 4: at most 4 pairs per tail tile
                                   (4100, 1088, 14336)])   # long K (FFN2), ragged n: 85 tiles = 1 wave + 11
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_streamk_tail_matches_unsplit(env, m, n, k, epi):
    """fgb_gemm_bf16_sk: the K range of the last, partly filled wave is spread over all CTA pairs. Same contract as the unsplit
    kernel (fp32 reference, 2e-3), near-identical to the unsplit result (summation order only), and the workspace flags are
    left clean: a second launch on the same workspace gives the same bits."""
    ops, o = env
    a, w, bias = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=1 / math.sqrt(k)), rnd(n, seed=3, scale=0.5)
    g0, g1, c0 = rnd(n, seed=4), rnd(n, seed=5), rnd(m, n, seed=6)
    rows0 = m // 3
    dev = torch.device("cuda", 0)
    ws = ops.gemm_workspace(dev)
    plain, out, again = c0.clone(), c0.clone(), c0.clone()
    ops.gemm(a, w, bias, plain, epi, g0, g1, rows0)
    ops.gemm_streamk_tune(dev, min_k=0)          # split short-K tiles too (the default only splits k >= 6144)
    try:
        ops.gemm(a, w, bias, out, epi, g0, g1, rows0, sk_ws=ws)
        ops.gemm(a, w, bias, again, epi, g0, g1, rows0, sk_ws=ws)
        ops.sync_check()
    finally:
        ops.gemm_streamk_tune(dev)
    y = (a.float() @ w.float().T + bias.float()).to(BF).float()
    if epi == 0:
        ref = y
    elif epi == 1:
        ref = F.gelu(y, approximate="tanh")
    elif epi == 2:
        gate = torch.where((torch.arange(m, device="cuda") < rows0)[:, None], g0.float()[None], g1.float()[None])
        ref = c0.float() + (gate * y).to(BF).float()
    else:
        ref = c0.float() + y
    assert rel_l2(out, ref.to(BF)) < 2e-3
    assert rel_l2(out, plain) < 1e-3
    assert torch.equal(out, again)
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0        # every flag lowered again


def test_gemm_no_bias_and_strided_views(env):
    ops, o = env
    a_wide = rnd(300, 1024, seed=1)
    a = a_wide[:, 256:768]                       # lda = 1024
    w = rnd(384, 512, seed=2, scale=0.05)
    out_wide = torch.zeros(300, 1024, dtype=BF, device="cuda")
    out = out_wide[:, 128:512]                   # ldc = 1024
    ops.gemm(a, w, None, out)
    ops.sync_check()
    assert rel_l2(out, (a.float() @ w.float().T).to(BF)) < 2e-3
    assert out_wide[:, :128].abs().max() == 0 and out_wide[:, 512:].abs().max() == 0   # nothing written outside


def test_gemm_rejects_bad_arguments(env):
    ops, o = env
    with pytest.raises(ValueError):
        ops.gemm(rnd(8, 64), rnd(16, 32), None, torch.empty(8, 16, dtype=BF, device="cuda"))
    with pytest.raises(RuntimeError, match="multiples of 8"):
        ops.gemm(rnd(8, 64), rnd(12, 64), None, torch.empty(8, 12, dtype=BF, device="cuda"))


@pytest.mark.parametrize("s_q,s_kv,heads", [(256, 128, 1), (128, 512, 2), (300, 200, 2), (48, 32, 2), (1000, 1000, 3),
                                            (30, 24, 2), (1025, 77, 4), (2304, 2304, 3)])
def test_attention_matches_sdpa(env, s_q, s_kv, heads):
    ops, o = env
    q, k, v = rnd(s_q, heads * 128, seed=1), rnd(s_kv, heads * 128, seed=2), rnd(s_kv, heads * 128, seed=3)
    out = torch.full((s_q, heads * 128), float("nan"), dtype=BF, device="cuda")
    ops.attention(q, k, v, out, heads)
    ops.sync_check()
    ref = o.attention(q[None].float(), k[None].float(), v[None].float(), heads)[0]
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 6e-3


def test_attention_strided_qkv_and_large_scores(env):
    """q|k|v as column slices of one fused buffer (the engine's layout); large-magnitude scores exercise the
    lazy running-max rescale of the accumulator."""
    ops, o = env
    heads, s = 2, 700
    qkv = rnd(s, 3 * heads * 128, seed=9)
    d = heads * 128
    qkv[:, :d] *= 6.0
    # make late keys dominate so the running max keeps growing
    qkv[:, d:2 * d] *= torch.linspace(0.2, 4.0, s, device="cuda")[:, None].to(BF)
    out = torch.empty(s, d, dtype=BF, device="cuda")
    ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, heads)
    ops.sync_check()
    ref = o.attention(qkv[None, :, :d].float(), qkv[None, :, d:2 * d].float(), qkv[None, :, 2 * d:].float(), heads)[0]
    assert rel_l2(out, ref) < 8e-3


def test_attention_properties_at_headline_size(env):
    """S = 27 280, 24 heads (BASELINE config 2): size-independent properties instead of an O(S^2) oracle.
    (1) rows of softmax sum to one: with V == 1 the output is exactly 1; (2) linearity in V;
    (3) a sample of query rows against the exact fp32 formula."""
    ops, o = env
    s, heads = 27280, 24
    d = heads * 128
    q, k = rnd(s, d, seed=1), rnd(s, d, seed=2)
    ones = torch.ones(s, d, dtype=BF, device="cuda")
    out = torch.empty(s, d, dtype=BF, device="cuda")
    ops.attention(q, k, ones, out, heads)
    ops.sync_check()
    assert (out.float() - 1).abs().max() < 1e-2
    v1, v2 = rnd(s, d, seed=3), rnd(s, d, seed=4)
    o1, o2, o12 = torch.empty_like(out), torch.empty_like(out), torch.empty_like(out)
    ops.attention(q, k, v1, o1, heads)
    ops.attention(q, k, v2, o2, heads)
    ops.attention(q, k, (v1.float() + v2.float()).to(BF), o12, heads)
    ops.sync_check()
    assert rel_l2(o12, o1.float() + o2.float()) < 2e-2
    rows = torch.tensor([0, 1, 127, 128, 255, 256, 13640, 27151, 27152, 27279], device="cuda")
    for h in (0, 11, 23):
        sl = slice(h * 128, (h + 1) * 128)
        sc = (q[rows, sl].float() @ k[:, sl].float().T) / math.sqrt(128)
        ref = torch.softmax(sc, dim=-1) @ v1[:, sl].float()
        assert rel_l2(o1[rows, sl], ref) < 1e-2


@pytest.mark.parametrize("s_q,s_kv,heads", [(128, 128, 1), (256, 64, 1), (300, 200, 2), (1000, 1000, 3), (70, 515, 2),
                                            (2304, 2304, 2), (5070, 512, 2)])
def test_attention_backward_matches_autograd(env, s_q, s_kv, heads):
    """fgb_attn_bwd vs torch autograd through the oracle's fp32 attention on the same bf16 inputs (config 5:
    the reference back-propagates through flash_attention, DIT:27-60)."""
    ops, o = env
    d = heads * 128
    q, k, v = rnd(s_q, d, seed=1), rnd(s_kv, d, seed=2), rnd(s_kv, d, seed=3)
    dout = rnd(s_q, d, seed=4)
    out = torch.empty(s_q, d, dtype=BF, device="cuda")
    lse = torch.full((heads, ops.stat_rows(s_q)), float("nan"), dtype=torch.float32, device="cuda")
    ops.attention(q, k, v, out, heads, lse=lse)
    dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    ops.attention_bwd(q, k, v, out, dout, lse, dq, dk, dv, heads)
    ops.sync_check()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = o.attention(qf[None], kf[None], vf[None], heads)[0]
    ref.backward(dout.float())
    # lse is in log2 units: log2(sum_j exp(s_ij))
    sc = torch.einsum("qhd,khd->hqk", q.float().view(s_q, heads, 128), k.float().view(s_kv, heads, 128)) / math.sqrt(128)
    assert (lse[:, :s_q] - torch.logsumexp(sc, dim=-1) * math.log2(math.e)).abs().max() < 2e-3
    assert torch.isfinite(lse).all()
    for got, want, name in ((dq, qf.grad, "dq"), (dk, kf.grad, "dk"), (dv, vf.grad, "dv")):
        assert torch.isfinite(got.float()).all(), name
        assert rel_l2(got, want) < 1e-2, (name, rel_l2(got, want))


def test_attention_backward_strided_fused_qkv(env):
    """q|k|v and dq|dk|dv as column slices of fused [S, 3*H*128] buffers (the training engine's layout)."""
    ops, o = env
    heads, s = 2, 450
    d = heads * 128
    qkv, dqkv = rnd(s, 3 * d, seed=5), torch.zeros(s, 3 * d, dtype=BF, device="cuda")
    dout, out = rnd(s, d, seed=6), torch.empty(s, d, dtype=BF, device="cuda")
    lse = torch.empty(heads, ops.stat_rows(s), dtype=torch.float32, device="cuda")
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    ops.attention(q, k, v, out, heads, lse=lse)
    ops.attention_bwd(q, k, v, out, dout, lse, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], heads)
    ops.sync_check()
    leaf = qkv.float().requires_grad_(True)
    o.attention(leaf[None, :, :d], leaf[None, :, d:2 * d], leaf[None, :, 2 * d:], heads)[0].backward(dout.float())
    assert rel_l2(dqkv, leaf.grad) < 1e-2


@pytest.mark.parametrize("rows,dim", [(30, 256), (1000, 3072), (129, 1024)])
def test_ln_modulate_and_affine(env, rows, dim):
    ops, o = env
    x = rnd(rows, dim, seed=1, scale=2.0) + 0.5
    sh0, sc0, sh1, sc1 = (rnd(dim, seed=s, scale=0.3) for s in (2, 3, 4, 5))
    n0 = rows // 4
    out = torch.empty_like(x)
    ops.ln_modulate(x, out, 1e-6, sh0, sc0, sh1, sc1, n0)
    ops.sync_check()
    first = (torch.arange(rows, device="cuda") < n0)[:, None]
    shift, scale = torch.where(first, sh0[None], sh1[None]), torch.where(first, sc0[None], sc1[None])
    ref_bf16 = o.modulate(o.layer_norm(x, 1e-6), shift, scale)                       # the reference's bf16 op chain
    ref_f32 = o.modulate(o.layer_norm(x.float(), 1e-6), shift.float(), scale.float())
    assert rel_l2(out, ref_f32) < 6e-3
    assert (out.float() - ref_bf16.float()).abs().max() <= 2 * ref_bf16.float().abs().max() * 2 ** -8
    assert (out == ref_bf16).float().mean() > 0.98   # same rounding points -> almost always bit-identical
    w, b = rnd(dim, seed=6, scale=0.2) + 1, rnd(dim, seed=7, scale=0.2)
    ops.ln_affine(x, out, 1e-6, w, b)
    ops.sync_check()
    assert rel_l2(out, o.layer_norm(x.float(), 1e-6, w.float(), b.float())) < 4e-3


def test_rmsnorm_rope_matches_reference_math(env, golden):
    ops, o = env
    cfg = o.TINY
    f, h, w = 2, 3, 5
    s = f * h * w
    tab = torch.from_numpy(ops.rope_table(128)).cuda()
    x = rnd(s, cfg.dim, seed=1, scale=1.5)
    wt = rnd(cfg.dim, seed=2, scale=0.1) + 1
    freqs = o.rope_freqs(o.rope_tables_3d(128), f, h, w)
    ref = o.rope_apply(o.rms_norm(x[None].cpu(), wt.cpu(), 1e-6), freqs, cfg.num_heads)[0]   # bf16 chain, fp64 rope
    got = x.clone()
    ops.rmsnorm_rope(got, 1e-6, wt, tab, (f, h, w), 0)
    ops.sync_check()
    assert rel_l2(got, ref.float()) < 3e-3
    assert (got.cpu() == ref).float().mean() > 0.97
    # no rope (cross-attention q/k) and a strided column slice of a wider buffer
    wide = rnd(s, 3 * cfg.dim, seed=3)
    ref2 = o.rms_norm(wide[:, cfg.dim:2 * cfg.dim].float(), wt.float(), 1e-6)
    keep = wide.clone()
    ops.rmsnorm_rope(wide[:, cfg.dim:2 * cfg.dim], 1e-6, wt)
    ops.sync_check()
    assert rel_l2(wide[:, cfg.dim:2 * cfg.dim], ref2) < 6e-3
    assert torch.equal(wide[:, :cfg.dim], keep[:, :cfg.dim]) and torch.equal(wide[:, 2 * cfg.dim:], keep[:, 2 * cfg.dim:])
    # sequence-parallel offsets: rank slices with ones-padding beyond the last token (USP:30-55)
    g = golden("usp")
    xs = torch.from_numpy(g["x"]).cuda().to(BF)
    ones = torch.ones(cfg.dim, dtype=BF, device="cuda")
    for rank in range(4):
        chunk = torch.from_numpy(g[f"chunk_rank{rank}"])[0].cuda().to(BF).contiguous()
        refr = o.sp_rope_apply(chunk[None].cpu(), freqs, cfg.num_heads, rank, 4)[0]
        # isolate the rotation: weight 1 and undo the RMS scale afterwards is awkward -> compare full op instead
        want = o.rope_apply(o.rms_norm(chunk[None].cpu(), ones.cpu(), 1e-6),
                            torch.cat([freqs, torch.ones(2, 1, 64, dtype=freqs.dtype)])[rank * 8:(rank + 1) * 8], cfg.num_heads)[0]
        ops.rmsnorm_rope(chunk, 1e-6, ones, tab, (f, h, w), rank * 8)
        ops.sync_check()
        assert rel_l2(chunk, want.float()) < 3e-3
        assert refr.shape == want.shape


def test_patchify_unpatchify_exact(env, golden):
    ops, o = env
    g = golden("ops")
    cfg = o.TINY
    lat = torch.from_numpy(g["lat"]).cuda().to(BF)[0].contiguous()          # [48, 2, 6, 10]
    grid = (2, 3, 5)
    rows = torch.full((30, 192), float("nan"), dtype=BF, device="cuda")
    ops.patchify_rows(lat, rows, grid, 0)
    ops.sync_check()
    ref = lat.view(48, 2, 3, 2, 5, 2).permute(1, 2, 4, 0, 3, 5).reshape(30, 192)   # (f h w) (c y z)
    assert torch.equal(rows, ref)
    # with a rank offset and zero padding past the last token
    part = torch.full((16, 192), float("nan"), dtype=BF, device="cuda")
    ops.patchify_rows(lat, part, grid, 16)
    ops.sync_check()
    assert torch.equal(part[:14], ref[16:]) and part[14:].abs().max() == 0
    # as a GEMM it equals the reference Conv3d patch embedding (golden from the real reference)
    w = o.make_weights(cfg, seed=0)
    out = torch.empty(30, cfg.dim, dtype=BF, device="cuda")
    ops.gemm(rows, w["patch_embedding.weight"].reshape(cfg.dim, -1).cuda().to(BF).contiguous(),
             w["patch_embedding.bias"].cuda().to(BF), out)
    ops.sync_check()
    want = torch.from_numpy(g["patchify"]).flatten(2).transpose(1, 2)[0]
    assert rel_l2(out, want) < 6e-3
    head_rows = rnd(30, 192, seed=5)
    got = torch.empty(48, 2, 6, 10, dtype=BF, device="cuda")
    ops.unpatchify(head_rows, got, grid)
    ops.sync_check()
    assert torch.equal(got, o.unpatchify(head_rows[None], grid, cfg)[0])


def test_cfg_fm_step_bit_exact_vs_reference(env, golden):
    ops, o = env
    from fairygen_b200.scheduler import FlowMatchScheduler
    g = golden("scheduler")
    s = FlowMatchScheduler("Wan")
    s.set_timesteps(50, shift=5.0)
    sample, npos, nneg, z0 = (torch.from_numpy(g[k]).cuda().to(BF) for k in ("sample", "npos", "nneg", "z0"))
    for i in (0, 17, 49):
        lat = sample.clone()
        s.step_fused(lat, npos, nneg, 5.0, i, z0)
        ops.sync_check()
        assert torch.equal(lat.float().cpu(), torch.from_numpy(g[f"step_{i}"])), f"step {i} differs from the reference"
    # reference-style API: step() returns a new tensor and leaves its inputs alone (FM:144-154)
    npred = (nneg + 5.0 * (npos - nneg))
    nxt = s.step(npred, s.timesteps[17], sample)
    nxt[:, :, 0:1] = z0
    assert torch.equal(nxt.float().cpu(), torch.from_numpy(g["step_17"]))
    assert torch.equal(sample.float().cpu(), torch.from_numpy(g["sample"]))
    # odd innermost size takes the scalar path
    lat = rnd(1, 3, 2, 5, 7, seed=1)
    a, b = rnd(1, 3, 2, 5, 7, seed=2), rnd(1, 3, 2, 5, 7, seed=3)
    want = lat + (b + 2.0 * (a - b)) * torch.tensor(s.sigma_delta(3))
    s.step_fused(lat, a, b, 2.0, 3, None)
    ops.sync_check()
    assert torch.equal(lat, want.to(BF))


def test_small_embedding_kernels(env, golden):
    ops, o = env
    g = golden("ops")
    t = torch.tensor([0.0, 996.0, 1000.0], device="cuda")
    out = torch.empty(3, 256, dtype=BF, device="cuda")
    ops.sinusoidal_embedding(t, out)
    ops.sync_check()
    assert (out.float().cpu() - torch.from_numpy(g["sinusoid_bf16"])).abs().max() <= 2 ** -8
    assert (out.float().cpu() == torch.from_numpy(g["sinusoid_bf16"])).float().mean() > 0.99
    x = rnd(7, 300, seed=1, scale=3)
    y = torch.empty_like(x)
    ops.silu(x, y)
    ops.sync_check()
    assert (y.float() - F.silu(x).float()).abs().max() <= 2 ** -7 * F.silu(x).float().abs().max()
    a, b = rnd(5, 600, seed=2), rnd(300, seed=3)
    outp = torch.empty_like(a)
    ops.add_bcast(a, b, outp, period=300)
    ops.sync_check()
    assert torch.equal(outp, a + torch.cat([b, b])[None])


def test_sp_pack_unpack_roundtrip(env):
    ops, o = env
    rows, heads, world = 37, 4, 2
    x = rnd(rows, 3 * heads * 128, seed=1)
    send = torch.empty(world * rows, 3 * heads * 128 // world, dtype=BF, device="cuda")
    ops.sp_pack_heads(x, send, heads, 3, world)
    ops.sync_check()
    ref = x.view(rows, 3, world, heads // world, 128).permute(2, 0, 1, 3, 4).reshape(world * rows, -1)
    assert torch.equal(send, ref)
    back = torch.zeros_like(x)
    ops.sp_unpack_heads(send, back, heads, 3, world)
    ops.sync_check()
    assert torch.equal(back, x)


@pytest.mark.parametrize("s_q,s_kv,heads,scale_q", [(300, 200, 2, 1.0), (1000, 1000, 3, 1.0), (2304, 2304, 2, 1.0), (515, 4100, 5, 1.0),
                                                    (700, 700, 2, 12.0)])
def test_bounded_score_attention(env, s_q, s_kv, heads, scale_q):
    """fgb_attn_fwd_bounded: softmax against the fixed Cauchy-Schwarz reference ||q_i||·max_j||k_j|| instead of a running
    maximum. scale_q = 12 pushes the bound past 60 (log2 units), which must fall back to the running-max path."""
    ops, o = env
    d = heads * 128
    q, k, v = rnd(s_q, d, seed=1, scale=scale_q), rnd(s_kv, d, seed=2), rnd(s_kv, d, seed=3)
    kmax2 = torch.empty(heads, dtype=torch.float32, device="cuda")
    ops.head_norm_max(k, kmax2, heads)
    ops.sync_check()
    want = (k.float().view(s_kv, heads, 128) ** 2).sum(-1).max(0).values
    assert torch.allclose(kmax2, want, rtol=1e-5)
    out = torch.full((s_q, d), float("nan"), dtype=BF, device="cuda")
    lse = torch.empty(heads, ops.stat_rows(s_q), dtype=torch.float32, device="cuda")
    ops.attention(q, k, v, out, heads, lse=lse, kmax2=kmax2)
    ops.sync_check()
    ref = o.attention(q[None].float(), k[None].float(), v[None].float(), heads)[0]
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 6e-3
    sc = torch.einsum("qhd,khd->hqk", q.float().view(s_q, heads, 128), k.float().view(s_kv, heads, 128)) / math.sqrt(128)
    assert (lse[:, :s_q] - torch.logsumexp(sc, dim=-1) * math.log2(math.e)).abs().max() < 5e-3 * max(1.0, scale_q)
    # odd head count (3 heads per rank under Ulysses SP8) and a strided k view
    kk = rnd(257, 3 * 128 + 64, seed=5)[:, 64:]
    km = torch.empty(3, dtype=torch.float32, device="cuda")
    ops.head_norm_max(kk, km, 3)
    assert torch.allclose(km, (kk.float().reshape(257, 3, 128) ** 2).sum(-1).max(0).values, rtol=1e-5)


@pytest.mark.parametrize("scale_q,want_mode", [(1.0, "bound_only"), (4.0, "bound_only"), (9.0, "first_tile_anchored"),
                                               (40.0, "running_max_fallback")])
def test_bounded_attention_reference_modes(env, scale_q, want_mode):
    """The three per-CTA modes of fgb_attn_fwd_bounded (AttnParams::kmax): bound <= 110 -> fixed reference from the bound alone;
    larger bounds -> anchored on the first tile's maximum; hopeless bounds -> running-max fallback.  All must give the same
    softmax, the CTA counters (fgb_attn_set_stats) must say which one ran, and a partial last KV tile must not leak
    probability mass (masked keys contribute exactly 0)."""
    ops, o = env
    heads, s_q, s_kv = 2, 600, 1000     # 1000 keys: the last KV tile has 104 live keys (wg1 of the last tile sees 40)
    d = heads * 128
    q, k, v = rnd(s_q, d, seed=1, scale=scale_q), rnd(s_kv, d, seed=2), rnd(s_kv, d, seed=3)
    kmax2 = torch.empty(heads, dtype=torch.float32, device="cuda")
    ops.head_norm_max(k, kmax2, heads)
    bound = float((q.float().view(s_q, heads, 128).norm(dim=-1).max() * kmax2.max().sqrt()) / math.sqrt(128) * math.log2(math.e))
    out = torch.full((s_q, d), float("nan"), dtype=BF, device="cuda")
    lse = torch.empty(heads, ops.stat_rows(s_q), dtype=torch.float32, device="cuda")
    ops.attention_stats_reset(q.device)
    ops.attention(q, k, v, out, heads, lse=lse, kmax2=kmax2)
    ops.sync_check()
    detail = ops.attention_stats_detail(q.device)
    print(f"scale_q {scale_q}: max bound {bound:.1f} log2 units, CTAs per mode {detail}")
    assert detail[want_mode] > 0 and sum(detail.values()) == heads * ((s_q + 255) // 256)
    ref = o.attention(q[None].float(), k[None].float(), v[None].float(), heads)[0]
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 6e-3
    sc = torch.einsum("qhd,khd->hqk", q.float().view(s_q, heads, 128), k.float().view(s_kv, heads, 128)) / math.sqrt(128)
    assert (lse[:, :s_q] - torch.logsumexp(sc, dim=-1) * math.log2(math.e)).abs().max() < 5e-3 * max(1.0, scale_q)
    # V == 1: every output must be exactly 1 (the row sum and the numerator see the same P, masked keys add nothing)
    ones = torch.ones_like(v)
    ops.attention(q, k, ones, out, heads, kmax2=kmax2)
    ops.sync_check()
    assert (out.float() - 1).abs().max() < 8e-3


@pytest.mark.parametrize("scale_q,want_mode", [(1.0, "bound_only"), (3.0, "bound_only"), (9.0, "first_tile_anchored"), (40.0, "running_max_fallback")])
@pytest.mark.parametrize("s_q,s_kv", [(600, 1000), (1300, 512)])
def test_head_level_query_bound(env, scale_q, want_mode, s_q, s_kv):
    """fgb_attn_fwd_bounded_qk: with qmax2[h] = max_i ||q_i||^2 a head whose bound sqrt(qmax2 kmax2)·scale·log2e <= 110 runs every
    row against ONE reference (no per-row norm pass, no CTA vote); heads above the window fall through to the per-row modes.
    Same softmax as the per-row form and as the fp32 reference; log-sum-exp consistent; V == 1 gives exactly 1. One head is made
    an outlier so that both branches run in one launch."""
    ops, o = env
    heads = 3
    d = heads * 128
    q, k, v = rnd(s_q, d, seed=1), rnd(s_kv, d, seed=2), rnd(s_kv, d, seed=3)
    q[:, 128:256] *= scale_q                    # head 1 carries the large queries; heads 0 and 2 stay small
    kmax2, qmax2 = (torch.empty(heads, dtype=torch.float32, device="cuda") for _ in range(2))
    ops.head_norm_max(k, kmax2, heads)
    ops.head_norm_max(q, qmax2, heads)
    out, per_row = (torch.full((s_q, d), float("nan"), dtype=BF, device="cuda") for _ in range(2))
    lse = torch.empty(heads, ops.stat_rows(s_q), dtype=torch.float32, device="cuda")
    ops.attention_stats_reset(q.device)
    ops.attention(q, k, v, out, heads, lse=lse, kmax2=kmax2, qmax2=qmax2)
    ops.sync_check()
    detail = ops.attention_stats_detail(q.device)
    items = (s_q + 255) // 256
    head_bound = (qmax2 * kmax2).sqrt() / math.sqrt(128) * math.log2(math.e)
    print(f"scale_q {scale_q}: head bounds {[round(float(b), 1) for b in head_bound]}, CTAs per mode {detail}")
    assert sum(detail.values()) == heads * items and detail[want_mode] >= items and detail["bound_only"] >= 2 * items
    ops.attention(q, k, v, per_row, heads, kmax2=kmax2)
    ops.sync_check()
    ref = o.attention(q[None].float(), k[None].float(), v[None].float(), heads)[0]
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 6e-3 and rel_l2(out, per_row) < 4e-3
    sc = torch.einsum("qhd,khd->hqk", q.float().view(s_q, heads, 128), k.float().view(s_kv, heads, 128)) / math.sqrt(128)
    assert (lse[:, :s_q] - torch.logsumexp(sc, dim=-1) * math.log2(math.e)).abs().max() < 5e-3 * max(1.0, scale_q)
    ops.attention(q, k, torch.ones_like(v), out, heads, kmax2=kmax2, qmax2=qmax2)
    ops.sync_check()
    assert (out.float() - 1).abs().max() < 8e-3


def test_norm_kernels_leave_the_query_bound(env):
    """qmax2 as a by-product: fgb_qk_norm_rope (per-warp and row-streaming path) and fgb_rmsnorm_hmax == fgb_head_norm_max of
    their own output, and the normalised rows are unchanged by asking for it."""
    ops, o = env
    import numpy as np
    tab = torch.from_numpy(np.ascontiguousarray(ops.rope_table(128))).cuda()
    for rows, dim in ((300, 768), (2500, 3072)):
        heads = dim // 128
        wq, wk = (1 + 0.1 * rnd(dim, seed=4).float()).to(BF), (1 + 0.1 * rnd(dim, seed=5).float()).to(BF)
        qkv = rnd(rows, 3 * dim, seed=3)
        a, b = qkv.clone(), qkv.clone()
        ka, kb, qa, want = (torch.full((heads,), -1.0, dtype=torch.float32, device="cuda") for _ in range(4))
        ops.qk_norm_rope(a, dim, 1e-6, wq, wk, tab, (5, 10, 10), 3, ka, qa)
        ops.qk_norm_rope(b, dim, 1e-6, wq, wk, tab, (5, 10, 10), 3, kb)
        ops.head_norm_max(a[:, :dim], want, heads)
        ops.sync_check()
        assert torch.equal(a, b) and torch.equal(ka, kb)
        assert_head_bound(qa, want)
        x = rnd(rows, dim, seed=6)
        y, z = x.clone(), x.clone()
        hm = torch.full((heads,), -1.0, dtype=torch.float32, device="cuda")
        ops.rmsnorm_rope(y, 1e-6, wq, hmax2=hm)
        ops.rmsnorm_rope(z, 1e-6, wq)
        ops.head_norm_max(y, want, heads)
        ops.sync_check()
        assert torch.equal(y, z)
        assert_head_bound(hm, want)


def test_attention_cta_pair_variant(env, monkeypatch):
    """The 2-CTA (tcgen05 cta_group::2) variant of the attention kernel, kept behind FGB_ATTN_PAIR=1 (it is slower on this
    workload, see csrc/attention.cu): same results as the default, including the key-split tail, ragged rows and all three
    reference modes."""
    ops, o = env
    for s_q, s_kv, heads, scale_q in ((600, 1000, 2, 1.0), (300, 4100, 5, 1.0), (1030, 700, 3, 9.0), (520, 260, 2, 40.0), (48, 32, 2, 1.0)):
        d = heads * 128
        q, k, v = rnd(s_q, d, seed=1, scale=scale_q), rnd(s_kv, d, seed=2), rnd(s_kv, d, seed=3)
        kmax2 = torch.empty(heads, dtype=torch.float32, device="cuda")
        ops.head_norm_max(k, kmax2, heads)
        outs = []
        for pair in ("0", "1"):
            monkeypatch.setenv("FGB_ATTN_PAIR", pair)
            out = torch.full((s_q, d), float("nan"), dtype=BF, device="cuda")
            lse = torch.empty(heads, ops.stat_rows(s_q), dtype=torch.float32, device="cuda")
            ops.attention(q, k, v, out, heads, lse=lse, kmax2=kmax2)
            plain = torch.full((s_q, d), float("nan"), dtype=BF, device="cuda")
            ops.attention(q, k, v, plain, heads)
            ops.sync_check()
            outs.append((out, lse[:, :s_q].clone(), plain))
        monkeypatch.delenv("FGB_ATTN_PAIR")
        ref = o.attention(q[None].float(), k[None].float(), v[None].float(), heads)[0]
        for out, lse, plain in outs:
            assert torch.isfinite(out.float()).all() and rel_l2(out, ref) < 6e-3 and rel_l2(plain, ref) < 8e-3
        assert rel_l2(outs[1][0], outs[0][0]) < 2e-3 and (outs[1][1] - outs[0][1]).abs().max() < 1e-2


@pytest.mark.parametrize("rows,dim,grid,tok0", [(300, 3072, (3, 10, 14), 0), (77, 768, (2, 6, 10), 13), (1000, 256, (10, 10, 10), 0)])
def test_qk_norm_rope_equals_the_three_separate_kernels(env, rows, dim, grid, tok0):
    """fgb_qk_norm_rope (one pass over the fused q|k|v rows) == fgb_rmsnorm_rope(q), fgb_rmsnorm_rope(k), fgb_head_norm_max(k),
    bit for bit on q and k, the same key bound; v is not touched."""
    ops, o = env
    import numpy as np
    heads = dim // 128
    qkv = rnd(rows, 3 * dim, seed=3)
    wq, wk = (1 + 0.1 * rnd(dim, seed=4).float()).to(BF), (1 + 0.1 * rnd(dim, seed=5).float()).to(BF)
    tab = torch.from_numpy(np.ascontiguousarray(ops.rope_table(128))).cuda()
    ref = qkv.clone()
    ops.rmsnorm_rope(ref[:, :dim], 1e-6, wq, tab, grid, tok0)
    ops.rmsnorm_rope(ref[:, dim:2 * dim], 1e-6, wk, tab, grid, tok0)
    kref = torch.empty(heads, dtype=torch.float32, device="cuda")
    ops.head_norm_max(ref[:, dim:2 * dim], kref, heads)
    got = qkv.clone()
    kmax = torch.full((heads,), -1.0, dtype=torch.float32, device="cuda")
    ops.qk_norm_rope(got, dim, 1e-6, wq, wk, tab, grid, tok0, kmax)
    ops.sync_check()
    assert torch.equal(got, ref)
    assert_head_bound(kmax, kref)


@pytest.mark.parametrize("rows,dim", [(5000, 3072), (2049, 768), (4097, 4096), (3000, 5120)])
def test_row_streaming_kernels_equal_the_per_warp_kernels(env, rows, dim):
    """Launches of >= 2048 rows take the persistent row-streaming kernels (bulk-copy ring + consumer warps); the arithmetic is
    the per-warp kernels', so the same rows computed in chunks of < 2048 rows (per-warp path) must give the same bits —
    ln_modulate (two modulation rows), ln_affine, rmsnorm (+RoPE, in place, strided view) and the fused q|k pass with its key bound."""
    ops, o = env
    import numpy as np
    chunk = 1500
    x = rnd(rows, dim, seed=1, scale=2.0) + 0.5
    sh0, sc0, sh1, sc1 = (rnd(dim, seed=s, scale=0.3) for s in (2, 3, 4, 5))
    n0 = rows // 3
    big, small = torch.empty_like(x), torch.empty_like(x)
    ops.ln_modulate(x, big, 1e-6, sh0, sc0, sh1, sc1, n0)
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        ops.ln_modulate(x[r0:r1], small[r0:r1], 1e-6, sh0, sc0, sh1, sc1, max(0, min(r1 - r0, n0 - r0)))
    assert torch.equal(big, small)
    ops.ln_affine(x, big, 1e-6, sc0, sh0)
    for r0 in range(0, rows, chunk):
        ops.ln_affine(x[r0:r0 + chunk], small[r0:r0 + chunk], 1e-6, sc0, sh0)
    assert torch.equal(big, small)
    # RMSNorm + RoPE in place on a strided column slice, ragged against the grid (rows beyond the grid are not rotated)
    heads = dim // 128
    grid = (5, 20, rows // 100 - 3)
    tab = torch.from_numpy(np.ascontiguousarray(ops.rope_table(128))).cuda()
    wq, wk = (1 + 0.1 * rnd(dim, seed=4).float()).to(BF), (1 + 0.1 * rnd(dim, seed=5).float()).to(BF)
    qkv = rnd(rows, 3 * dim, seed=3)
    a, b = qkv.clone(), qkv.clone()
    ops.rmsnorm_rope(a[:, :dim], 1e-6, wq, tab, grid, 7)
    ops.rmsnorm_rope(a[:, dim:2 * dim], 1e-6, wk)
    for r0 in range(0, rows, chunk):
        ops.rmsnorm_rope(b[r0:r0 + chunk, :dim], 1e-6, wq, tab, grid, 7 + r0)
        ops.rmsnorm_rope(b[r0:r0 + chunk, dim:2 * dim], 1e-6, wk)
    assert torch.equal(a, b)
    a, b = qkv.clone(), qkv.clone()
    ka = torch.full((heads,), -1.0, dtype=torch.float32, device="cuda")
    ops.qk_norm_rope(a, dim, 1e-6, wq, wk, tab, grid, 7, ka)
    kb = torch.zeros(heads, dtype=torch.float32, device="cuda")
    for r0 in range(0, rows, chunk):
        kc = torch.empty(heads, dtype=torch.float32, device="cuda")
        ops.qk_norm_rope(b[r0:r0 + chunk], dim, 1e-6, wq, wk, tab, grid, 7 + r0, kc)
        kb = torch.maximum(kb, kc)
    ops.sync_check()
    assert torch.equal(a, b) and torch.equal(a[:, 2 * dim:], qkv[:, 2 * dim:])
    assert torch.allclose(ka, kb, rtol=1e-6)


def test_exchange_barrier_reports_a_missing_peer(env):
    """fgb_sp_barrier_status on ONE GPU with a stand-in peer (a second flag array on the same device that nobody drives): the
    barrier publishes its epoch to the peer, gives up after the requested number of clocks, writes the epoch to the status word
    and returns (the reference's NCCL all-to-all, USP:136-141, would hang; VERDICT r1 asked for a status instead of a trap).
    Once the "peer" has published, the same call passes and leaves the status alone."""
    ops, _ = env
    dev = torch.device("cuda", 0)
    mine = torch.zeros(64, dtype=torch.int32, device=dev)
    peer = torch.zeros(64, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ptrs = [mine.data_ptr(), peer.data_ptr()]
    ops.sp_barrier(dev, ptrs, 2, 0, 5, status=status, timeout_clocks=3_000_000)       # ~2 ms
    ops.sync_check()
    assert int(status.item()) == 5
    assert int(peer[0].item()) == 5 and int(mine[0].item()) == 5 and int(mine[1].item()) == 0
    # an exchange that has lost a peer does not wait out every later barrier (60 per forward): with the status word set the
    # default ~30 s wait ends within ~1000 polls, and the FIRST epoch that timed out stays in the word
    import time
    t0 = time.perf_counter()
    ops.sp_barrier(dev, ptrs, 2, 0, 6, status=status)
    ops.sync_check()
    assert time.perf_counter() - t0 < 2.0
    assert int(status.item()) == 5 and int(peer[0].item()) == 6
    status.zero_()
    mine[1] = 8                                    # the peer is already one exchange ahead: epochs only have to be reached
    ops.sp_barrier(dev, ptrs, 2, 0, 7, status=status, timeout_clocks=3_000_000)
    ops.sync_check()
    assert int(status.item()) == 0 and int(peer[0].item()) == 7
    with pytest.raises(ValueError):
        ops.sp_barrier(dev, ptrs, 2, 0, 8, status=torch.zeros(1, device=dev))
    with pytest.raises(RuntimeError):
        ops.sp_barrier(dev, ptrs, 9, 0, 8, status=status)
