"""Host logic of the DiT engine (fairygen_b200/engine.py) on the CPU: every kernel is replaced by a plain-torch statement of its
contract (include/fairygen_b200.h), and the orchestration — two-row timestep tables, which rows take the t = 0 modulation,
fused q|k|v weights, RoPE table / grid / token offset, context K|V cache, patchify / unpatchify layouts — must reproduce the
pinned oracle and the reference's stored forwards.  The kernels themselves are covered by the `-m gpu` tests."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import wan_dit_oracle as o

BF = torch.bfloat16
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_forward.npz"))


def _emulated_ops(monkeypatch):
    from fairygen_b200 import ops

    def ln(x, eps):
        xf = x.float()
        return (xf - xf.mean(-1, keepdim=True)) * torch.rsqrt(xf.var(-1, unbiased=False, keepdim=True) + eps)

    def gemm_workspace(device):
        return torch.zeros(16, dtype=torch.uint8)

    def gemm(a, w, bias, out, epilogue=ops.EPI_BIAS, gate0=None, gate1=None, rows_gate0=0, a2=None, w2=None, sk_ws=None):
        acc = a.float() @ w.float().T
        if a2 is not None:
            acc = acc + a2.float() @ w2.float().T
        y = (acc + (0 if bias is None else bias.float())).to(BF).float()
        if epilogue == ops.EPI_BIAS_GELU_TANH:
            y = F.gelu(y, approximate="tanh")
        elif epilogue == ops.EPI_GATED_RESIDUAL:
            first = (torch.arange(a.shape[0]) < rows_gate0)[:, None]
            gate = torch.where(first, gate0.float()[None], gate1.float()[None])
            y = out.float() + (gate * y).to(BF).float()
        elif epilogue == ops.EPI_RESIDUAL:
            y = out.float() + y
        out.copy_(y.to(BF))
        return out

    def sinusoidal_embedding(ts, out):
        out.copy_(o.sinusoidal_embedding_1d(out.shape[1], ts).to(BF))
        return out

    def silu(x, out):
        out.copy_(F.silu(x.float()).to(BF))
        return out

    def add_bcast(a, b, out, period=None):
        rows, cols = a.shape
        period = cols if period is None else period
        out.copy_((a.float() + b.reshape(-1)[:period].float().repeat(cols // period)[None]).to(BF))
        return out

    def ln_modulate(x, out, eps, shift0, scale0, shift1, scale1, rows_mod0):
        first = (torch.arange(x.shape[0]) < rows_mod0)[:, None]
        shift = torch.where(first, shift0.float()[None], shift1.float()[None])
        scale = torch.where(first, scale0.float()[None], scale1.float()[None])
        out.copy_((ln(x, eps).to(BF).float() * (1 + scale) + shift).to(BF))
        return out

    def ln_affine(x, out, eps, weight, bias):
        out.copy_((ln(x, eps) * weight.float() + bias.float()).to(BF))
        return out

    def rmsnorm_rope(x, eps, weight, rope_tab=None, grid=(1, 1, 1), token_offset=0, hmax2=None):
        xf = x.float()
        y = ((xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(BF).float() * weight.float()).to(BF).float()
        if rope_tab is not None:
            f, h, w = grid
            rows, dim = x.shape
            t = torch.arange(rows) + token_offset
            pos = torch.stack([t // (h * w), (t // w) % h, t % w], 1)                         # [rows, 3]
            lanes = torch.tensor([0] * 22 + [1] * 21 + [2] * 21)                              # frame / row / column lanes of a head
            tab = rope_tab[pos[:, lanes], torch.arange(64)]                                    # [rows, 64, 2] (cos, sin)
            z = y.view(rows, dim // 128, 64, 2)
            re = z[..., 0] * tab[:, None, :, 0] - z[..., 1] * tab[:, None, :, 1]
            im = z[..., 0] * tab[:, None, :, 1] + z[..., 1] * tab[:, None, :, 0]
            y = torch.stack([re, im], -1).reshape(rows, dim)
        x.copy_(y.to(BF))
        if hmax2 is not None:
            hmax2.copy_(x.float().view(x.shape[0], -1, 128).pow(2).sum(-1).max(0).values)
        return x

    def head_norm_max(k, out_f32, heads):
        out_f32.copy_(k.float().view(k.shape[0], heads, 128).pow(2).sum(-1).max(0).values)
        return out_f32

    def qk_norm_rope(qkv, dim, eps, wq, wk, rope_tab, grid, token_offset, kmax2, qmax2=None):
        rmsnorm_rope(qkv[:, :dim], eps, wq, rope_tab, grid, token_offset)
        rmsnorm_rope(qkv[:, dim:2 * dim], eps, wk, rope_tab, grid, token_offset)
        head_norm_max(qkv[:, dim:2 * dim], kmax2, dim // 128)
        if qmax2 is not None:
            head_norm_max(qkv[:, :dim], qmax2, dim // 128)

    def attention(q, k, v, out, heads, scale=None, lse=None, kmax2=None, qmax2=None):
        qf, kf, vf = (t.float().view(t.shape[0], heads, 128).transpose(0, 1) for t in (q, k, v))
        p = torch.softmax(qf @ kf.transpose(1, 2) / 128 ** 0.5, -1)
        out.copy_((p @ vf).transpose(0, 1).reshape(q.shape[0], heads * 128).to(BF))
        return out

    def patchify_rows(latents, rows_out, grid, token_offset=0):
        C = latents.shape[0]
        f, h, w = grid
        x = latents.view(C, f, h, 2, w, 2).permute(1, 2, 4, 0, 3, 5).reshape(f * h * w, C * 4)     # row[t, c*4 + y*2 + z]
        rows_out.zero_()
        n = max(0, min(rows_out.shape[0], f * h * w - token_offset))
        rows_out[:n, :C * 4] = x[token_offset:token_offset + n]
        return rows_out

    def unpatchify(head_rows, out, grid):
        C = out.shape[0]
        f, h, w = grid
        x = head_rows[:f * h * w, :4 * C].view(f, h, w, 2, 2, C)                                   # rows[t, y*2C + z*C + c]
        out.copy_(x.permute(5, 0, 1, 3, 2, 4).reshape(C, f, 2 * h, 2 * w))
        return out

    def cfg_fm_step(latents, noise_pos, noise_neg, first_frame, cfg_scale, sigma_delta):
        n = noise_pos if noise_neg is None else (noise_neg.float() + (cfg_scale * (noise_pos.float() - noise_neg.float()).to(BF).float()).to(BF).float()).to(BF)
        latents.copy_((latents.float() + (n.float().view_as(latents) * sigma_delta).to(BF).float()).to(BF))
        if first_frame is not None:
            latents[..., 0:1, :, :] = first_frame.view(latents[..., 0:1, :, :].shape)
        return latents

    for name, fn in list(locals().items()):
        if callable(fn) and name != "ln" and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(ops, "context", lambda device: None)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.fixture()
def engine(monkeypatch):
    _emulated_ops(monkeypatch)
    import fairygen_b200 as fg
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w = o.make_weights(o.TINY, seed=0)
    from fairygen_b200 import ops
    eng = fg.WanDiTEngine.__new__(fg.WanDiTEngine)        # the constructor refuses non-CUDA devices (no CPU path in the product);
    eng.cfg, eng.device, eng.ctx, eng.sp = cfg, torch.device("cpu"), None, None     # the test sets up the same fields by hand
    eng.rope_tab = torch.from_numpy(ops.rope_table(cfg.head_dim))
    eng._init_state()
    eng.load_state_dict(w)
    return eng, {k: v.to(BF).float() for k, v in w.items()}


@pytest.mark.parametrize("ts_val,fused,key", [(900.0, True, "fused_t900"), (37.0, True, "fused_t37"), (900.0, False, "plain_t900")])
def test_forward_orchestration_matches_oracle_and_reference(engine, ts_val, fused, key):
    eng, w16 = engine
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 8, 8), text_len=32, live_text=8)
    ctx = cp if fused else cn
    ts = torch.tensor([ts_val])
    out = eng.forward(lat.to(BF), ts, ctx.to(BF), fused)
    with torch.no_grad():
        want = o.dit_forward(w16, o.TINY, lat.to(BF).float(), ts, ctx.to(BF).float(), fused)
    assert out.shape == lat.shape
    assert rel(out.float(), want) < 1e-2, rel(out.float(), want)
    assert rel(out.float(), torch.from_numpy(GOLD[key])) < 1.5e-2            # the reference model's own fp32 output
    again = eng.forward(lat.to(BF), ts, ctx.to(BF), fused)                    # second call: context K|V from the cache
    assert torch.equal(out, again)


def test_ragged_grid(engine):
    eng, w16 = engine
    lat, _, cp, _ = o.make_inputs(o.TINY, (1, 48, 2, 6, 10), text_len=24, live_text=8)
    ts = torch.tensor([500.0])
    out = eng.forward(lat.to(BF), ts, cp.to(BF), True)
    with torch.no_grad():
        want = o.dit_forward(w16, o.TINY, lat.to(BF).float(), ts, cp.to(BF).float(), True)
    assert rel(out.float(), want) < 1e-2


def test_the_real_constructor_still_refuses_the_cpu():
    import fairygen_b200 as fg
    with pytest.raises(RuntimeError):
        fg.WanDiTEngine(fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2), "cpu")


def test_four_step_cfg_denoise_matches_the_reference_loop(engine, monkeypatch):
    """WanDenoiser (2 forwards + the fused CFG / Euler / first-frame step per iteration, PIPE:285-309) on the emulated kernels
    against the reference's own 4-step loop (tests/golden/denoise.npz)."""
    import fairygen_b200 as fg
    from fairygen_b200 import ops, scheduler

    def step_fused(self, latents, noise_pos, noise_neg, cfg_scale, index, first_frame_latents=None, to_final=False):
        # FlowMatchScheduler.step_fused without its CUDA-only guard (tests/test_host_logic.py checks the guard itself)
        ops.cfg_fm_step(latents, noise_pos, noise_neg, first_frame_latents, float(cfg_scale), self.sigma_delta(index, to_final))
        return latents

    monkeypatch.setattr(scheduler.FlowMatchScheduler, "step_fused", step_fused)
    eng, _ = engine
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 8, 8), text_len=32, live_text=8)
    out = fg.WanDenoiser(eng, num_inference_steps=4, cfg_scale=5.0, sigma_shift=5.0)(lat, cp, cn, z0)
    gold = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "golden", "denoise.npz"))["final"])
    assert torch.equal(out[:, :, 0:1], z0.to(BF))
    assert rel(out.float(), gold) < 3e-2, rel(out.float(), gold)            # north star: <= 3e-2 after a schedule


def test_query_bound_plumbing_does_not_change_the_forward(engine):
    """FGB_QMAX / engine.query_bounds: the norm kernels are asked for the per-head query bounds and the attention calls receive
    them (fgb_attn_fwd_bounded_qk) — the result is the same softmax, so the forward must not move."""
    eng, w16 = engine
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 8, 8), text_len=32, live_text=8)
    ts = torch.tensor([500.0])
    base = eng.forward(lat.to(BF), ts, cp.to(BF), True)
    seen = []
    from fairygen_b200 import ops
    inner = ops.attention

    def spy(q, k, v, out, heads, scale=None, lse=None, kmax2=None, qmax2=None):
        seen.append(None if qmax2 is None else qmax2.clone())
        return inner(q, k, v, out, heads, scale=scale, lse=lse, kmax2=kmax2, qmax2=qmax2)

    ops.attention = spy
    try:
        for bits, want in ((0, [False, False]), (1, [True, False]), (2, [False, True]), (3, [True, True])):
            eng.query_bounds = bits
            seen.clear()
            out = eng.forward(lat.to(BF), ts, cp.to(BF), True)
            assert torch.equal(out, base)
            per_block = [(s is not None) for s in seen[:2]]          # block 0: self-attention call, cross-attention call
            assert per_block == want, (bits, per_block)
            for s in seen:
                assert s is None or (s.shape == (eng.cfg.num_heads,) and bool((s > 0).all()))
    finally:
        ops.attention = inner
        eng.query_bounds = 0


def test_denoiser_uploads_host_prompt_embeddings_once():
    """WanDenoiser.step accepts the prompt embeddings as HOST tensors on every step (what a pipeline that keeps them on the CPU
    does): equal content resolves to the same device tensor (so the engine's context cache hits by identity), new content
    uploads once, device tensors pass through, and only the last four contexts are remembered."""
    import types

    from fairygen_b200.pipeline import WanDenoiser

    den = WanDenoiser.__new__(WanDenoiser)
    den.engine = types.SimpleNamespace(device=torch.device("cpu"))
    den._host_contexts = []
    a = torch.randn(1, 16, 8)
    d1 = den._context_on_device(a)
    assert d1.dtype == BF and den._context_on_device(a) is d1 and den._context_on_device(a.clone()) is d1
    b = a.clone()
    b[0, 0, 0] += 1
    d2 = den._context_on_device(b)
    assert d2 is not d1 and den._context_on_device(b.clone()) is d2 and den._context_on_device(a) is d1
    a.add_(1)                                   # the caller edits its tensor in place: the private host copy still says "old content"
    assert den._context_on_device(a) is not d1
    assert den._context_on_device(None) is None
    for i in range(6):
        den._context_on_device(torch.full((1, 4, 8), float(i)))
    assert len(den._host_contexts) == 4
