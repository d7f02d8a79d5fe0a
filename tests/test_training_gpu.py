"""Stage-2 LoRA fine-tune step (BASELINE config 5) on the B200: every training kernel against torch autograd on the
same bf16 inputs, then the whole step (loss, prediction, all lora_B2 gradients) against the training oracle, which
tests/test_train_oracle.py pins to the REAL reference DiT under autograd (tests/golden/train.npz)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def env():
    import fairygen_b200
    from fairygen_b200 import ops
    from oracle import wan_dit_oracle as o
    from oracle import wan_train_oracle as t
    torch.cuda.set_device(0)
    ops.context(torch.device("cuda", 0))
    return fairygen_b200, ops, o, t


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


# ---------------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n_in,k_out", [(128, 256, 64), (200, 3072, 192), (1000, 3072, 3072), (513, 256, 1536), (320, 3072, 9216),
                                          (300, 14336, 3072), (2049, 3072, 14336)])
def test_gemm_dgrad(env, m, n_in, k_out):
    _, ops, _, _ = env
    dy, w = rnd(m, k_out, seed=1), rnd(k_out, n_in, seed=2, scale=1 / math.sqrt(k_out))
    dx = torch.full((m, n_in), float("nan"), dtype=BF, device="cuda")
    ops.gemm_dgrad(dy, w, dx)
    ops.sync_check()
    assert rel_l2(dx, dy.float() @ w.float()) < 3e-3


@pytest.mark.parametrize("rows,dim", [(30, 256), (1000, 3072), (129, 1024)])
def test_ln_backward(env, rows, dim):
    _, ops, o, _ = env
    x, dy, dres = rnd(rows, dim, seed=1, scale=2.0), rnd(rows, dim, seed=2), rnd(rows, dim, seed=3)
    sc0, sc1, wgt, bias = rnd(dim, seed=4, scale=0.3), rnd(dim, seed=5, scale=0.3), 1 + rnd(dim, seed=6, scale=0.1), rnd(dim, seed=7)
    n0 = rows // 3
    # modulate variant: y = LN(x) * (1 + scale[row]) + shift
    xf = x.float().requires_grad_(True)
    scale_rows = torch.where((torch.arange(rows, device="cuda") < n0)[:, None], sc0.float()[None], sc1.float()[None])
    (o.layer_norm(xf, 1e-6) * (1 + scale_rows)).backward(dy.float())
    out = dres.clone()
    ops.ln_bwd(x, dy, out, 1e-6, sc0, sc1, n0, affine=False, dres=out)
    ops.sync_check()
    assert rel_l2(out, dres.float() + xf.grad) < 6e-3
    out2 = torch.empty_like(x)
    ops.ln_bwd(x, dy, out2, 1e-6, sc0, sc1, n0, affine=False, dres=None)
    assert rel_l2(out2, xf.grad) < 6e-3
    # affine variant (norm3)
    xf = x.float().requires_grad_(True)
    o.layer_norm(xf, 1e-6, wgt.float(), bias.float()).backward(dy.float())
    out3 = dres.clone()
    ops.ln_bwd(x, dy, out3, 1e-6, wgt, None, 0, affine=True, dres=out3)
    ops.sync_check()
    assert rel_l2(out3, dres.float() + xf.grad) < 6e-3


@pytest.mark.parametrize("grid,dim,heads", [((2, 3, 5), 256, 2), ((3, 4, 4), 3072, 24), ((1, 1, 7), 1024, 8)])
def test_rmsnorm_rope_backward(env, grid, dim, heads):
    _, ops, o, _ = env
    rows = grid[0] * grid[1] * grid[2]
    x, dy, w = rnd(rows, 3 * dim, seed=1, scale=1.5), rnd(rows, dim, seed=2), 1 + rnd(dim, seed=3, scale=0.1)
    x_pre = x[:, dim:2 * dim]                      # strided view, like k inside the fused qkv buffer
    tab = torch.from_numpy(ops.rope_table(128)).cuda()
    freqs = o.rope_freqs(o.rope_tables_3d(128), *grid).cuda()
    xf = x_pre.float().requires_grad_(True)
    o.rope_apply(o.rms_norm(xf[None], w.float(), 1e-6), freqs, heads)[0].float().backward(dy.float())
    d = dy.clone()
    ops.rmsnorm_rope_bwd(x_pre, d, 1e-6, w, tab, grid, 0)
    ops.sync_check()
    assert rel_l2(d, xf.grad) < 6e-3
    # without RoPE (cross-attention q / k)
    xf = x_pre.float().requires_grad_(True)
    o.rms_norm(xf[None], w.float(), 1e-6)[0].backward(dy.float())
    d = dy.clone()
    ops.rmsnorm_rope_bwd(x_pre, d, 1e-6, w)
    ops.sync_check()
    assert rel_l2(d, xf.grad) < 6e-3


def test_gelu_gate_loss_noise_unpatchify_adamw(env):
    _, ops, o, _ = env
    z, dh = rnd(1000, 512, seed=1, scale=2.0), rnd(1000, 512, seed=2)
    h, dz = torch.empty_like(z), torch.empty_like(z)
    ops.gelu_tanh(z, h)
    ops.gelu_tanh_bwd(z, dh, dz)
    zf = z.float().requires_grad_(True)
    ref = F.gelu(zf, approximate="tanh")
    ref.backward(dh.float())
    assert rel_l2(h, ref) < 4e-3 and rel_l2(dz, zf.grad) < 5e-3
    # gate
    dx, g0, g1 = rnd(77, 256, seed=3), rnd(256, seed=4), rnd(256, seed=5)
    out = torch.empty_like(dx)
    ops.mul_gate(dx, out, g0, g1, 20)
    want = dx.float() * torch.where((torch.arange(77, device="cuda") < 20)[:, None], g0.float()[None], g1.float()[None])
    assert torch.equal(out, want.to(BF))
    # add_noise / training_target with torch's bf16 op chain (flow_match.py:164-175)
    x0, nz = rnd(1, 48, 3, 8, 8, seed=6), rnd(1, 48, 3, 8, 8, seed=7)
    sigma = torch.tensor(0.8333333, dtype=torch.float32)
    lat, tgt = torch.empty_like(x0), torch.empty_like(x0)
    ops.fm_noise_target(x0, nz, float(sigma), lat, tgt)
    assert torch.equal(lat, (1 - sigma) * x0 + sigma * nz) and torch.equal(tgt, nz - x0)
    # loss + gradient (loss.py:19-20)
    pred = rnd(1, 48, 3, 8, 8, seed=8)
    pf = pred.clone().requires_grad_(True)
    loss_ref = F.mse_loss(pf.float(), tgt.float()) * 1.7
    loss_ref.backward()
    loss, dpred = torch.zeros(1, device="cuda"), torch.empty_like(pred)
    ops.mse_loss_grad(pred, tgt, 1.7, loss, dpred)
    ops.sync_check()
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * float(loss_ref) + 1e-7
    assert rel_l2(dpred, pf.grad) < 4e-3
    # unpatchify adjoint
    grid, ch = (2, 3, 5), 48
    dp = rnd(ch, 2, 6, 10, seed=9)
    rows = torch.zeros(30, 4 * ch, dtype=BF, device="cuda")
    ops.unpatchify_bwd(dp, rows, grid)
    hr = torch.randn(1, 30, 4 * ch, device="cuda").requires_grad_(True)
    o.unpatchify(hr, grid, o.TINY).backward(dp.float()[None])
    assert torch.equal(rows.float(), hr.grad[0].to(BF).float())
    # AdamW
    p, g = rnd(5000, seed=10), torch.randn(5000, device="cuda")
    m, v = torch.zeros(5000, device="cuda"), torch.zeros(5000, device="cuda")
    pr = torch.nn.Parameter(p.float().clone())
    opt = torch.optim.AdamW([pr], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    q = p.clone()
    for step in (1, 2):
        pr.grad = g.clone()
        opt.step()
        ops.adamw_step(q, g, m, v, 1e-2, 0.9, 0.999, 1e-8, 1e-2, step)
        pr.data = pr.data.to(BF).float()     # the kernel keeps bf16 parameters
    ops.sync_check()
    assert rel_l2(q, pr.data) < 4e-3


@pytest.mark.parametrize("n,k,rank", [(256, 256, 32), (3072, 3072, 32), (512, 256, 16), (14336, 3072, 32)])
def test_lora_merge_and_wgrad(env, n, k, rank):
    _, ops, _, _ = env
    w, a1 = rnd(n, k, seed=1, scale=1 / math.sqrt(k)), rnd(rank, k, seed=2, scale=1 / math.sqrt(k))
    b1, b2 = rnd(n, rank, seed=3, scale=0.02), rnd(n, rank, seed=4, scale=0.02)
    mask = (torch.rand(n, rank, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) > 0.5).to(torch.uint8)
    w_eff = torch.empty_like(w)
    ops.lora_merge(w, a1, b1, b2, mask, w_eff, 2.0, 1.0)
    ops.sync_check()
    want = w.float() + (b1.float() + b2.float() * mask.float() * 2.0) @ a1.float()
    assert rel_l2(w_eff, want) < 3e-3
    # merged forward == TMOD:317-352 evaluated op by op
    x = rnd(300, k, seed=6)
    y_ref = x.float() @ w.float().T + (x.float() @ a1.float().T) @ b1.float().T + (x.float() @ a1.float().T) @ (b2.float() * mask.float() * 2).T
    assert rel_l2(x.float() @ w_eff.float().T, y_ref) < 3e-3
    # wgrad: dB2 = (dy^T t) * mask * 2, strided dy / t views, accumulation on top of existing content
    rows = 777
    dy_full, t_full = rnd(rows, n + 64, seed=7), rnd(rows, 3 * rank, seed=8)
    dy, t = dy_full[:, 64:], t_full[:, rank:2 * rank]
    db = torch.ones(n, rank, dtype=torch.float32, device="cuda")
    ops.lora_wgrad(dy, t, db, mask, 2.0)
    ops.sync_check()
    want = 1.0 + (dy.float().T @ t.float()) * mask.float() * 2.0
    assert rel_l2(db, want) < 1e-3
    keep = torch.empty(n, rank, dtype=torch.uint8, device="cuda")
    ops.bernoulli_mask(keep, 0.5, 1234)
    assert 0.47 < float(keep.float().mean()) < 0.53
    keep2 = torch.empty_like(keep)
    ops.bernoulli_mask(keep2, 0.5, 1234)
    assert torch.equal(keep, keep2)


# ---------------------------------------------------------------------------------------------------------------
# whole step
# ---------------------------------------------------------------------------------------------------------------
def _setup(fg, o, t, cfg, ocfg, rank=32):
    from fairygen_b200.training import Stage2Trainer
    w = o.make_weights(ocfg, seed=0)
    lora = o.make_lora(ocfg, rank=rank, seed=2)
    b2, masks = t.make_b2(ocfg, rank=rank), t.make_masks(ocfg, rank=rank)
    eng = fg.WanDiTEngine(cfg, "cuda")
    eng.load_state_dict(w)
    return w, lora, b2, masks, eng, Stage2Trainer


def _oracle_on_gpu(o, t, ocfg, w, lora, b2, masks, x0, noise, timestep_id, ctx):
    """fp32 oracle on the GPU, fed the bf16-rounded tensors the CUDA path sees (weights, adapters, inputs)."""
    r = lambda v: v.to(BF).float().cuda()  # noqa: E731
    return t.loss_and_grads({k: r(v) for k, v in w.items()}, ocfg, {k: r(v) for k, v in lora.items()}, {k: r(v) for k, v in b2.items()},
                            {k: v.cuda() for k, v in masks.items()}, r(x0), r(noise), timestep_id, r(ctx), timestep_dtype=BF)


@pytest.mark.parametrize("shape,text_len,timestep_id", [((1, 48, 3, 8, 8), 32, 500), ((1, 48, 2, 6, 10), 24, 37)])
@pytest.mark.parametrize("recompute", [False, True])
def test_tiny_training_step_vs_oracle(env, shape, text_len, timestep_id, recompute):
    fg, ops, o, t = env
    ocfg = o.TINY
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w, lora, b2, masks, eng, Stage2Trainer = _setup(fg, o, t, cfg, ocfg)
    x0, _, ctx, _ = o.make_inputs(ocfg, shape, text_len=text_len, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    tr = Stage2Trainer(eng, lora, rank=32, recompute=recompute)
    tr.load_b2(b2)
    tr.zero_grad()
    loss, pred = tr.step(x0, noise, timestep_id, ctx.cuda(), masks=masks, return_pred=True)
    ops.sync_check()
    loss_ref, pred_ref, grads_ref = _oracle_on_gpu(o, t, ocfg, w, lora, b2, masks, x0, noise, timestep_id, ctx)
    assert rel_l2(pred, pred_ref) < 1e-2                      # north star: per-forward latent rel-L2
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    worst = 0.0
    for name in tr.targets:
        g, ref = tr.grad[name], grads_ref[name]
        assert torch.isfinite(g).all(), name
        assert torch.all(g[masks[name].cuda() == 0] == 0), name       # dropped entries get no gradient
        worst = max(worst, rel_l2(g, ref))
        assert rel_l2(g, ref) < 5e-2, (name, rel_l2(g, ref))
    print(f"tiny step {shape} recompute={recompute}: pred {rel_l2(pred, pred_ref):.3e} loss {float(loss):.6f} vs {float(loss_ref):.6f} "
          f"worst grad rel-L2 {worst:.3e}")
    # gradients accumulate across micro-steps like autograd's .grad
    before = tr.grad_flat.clone()
    tr.step(x0, noise, timestep_id, ctx.cuda(), masks=masks)
    ops.sync_check()
    assert rel_l2(tr.grad_flat, 2 * before) < 1e-3


def test_autograd_bridge_matches_step_and_optimizer_moves_b2(env):
    """pipe.model_fn under autograd (loss.py:17): torch computes the loss, our Function the backward."""
    fg, ops, o, t = env
    ocfg = o.TINY
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w, lora, b2, masks, eng, Stage2Trainer = _setup(fg, o, t, cfg, ocfg)
    shape = (1, 48, 3, 8, 8)
    x0, _, ctx, _ = o.make_inputs(ocfg, shape, text_len=32, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    tr = Stage2Trainer(eng, lora, rank=32)
    tr.load_b2(b2)
    tr.zero_grad()
    tr.step(x0, noise, 500, ctx.cuda(), masks=masks)
    want = tr.grad_flat.clone()
    params = [torch.nn.Parameter(b2[n].to(BF).cuda()) for n in tr.targets]
    sched = tr.scheduler
    timestep = sched.timesteps[500:501].to(BF).cuda()
    x0b, nzb = x0.to(BF).cuda(), noise.to(BF).cuda()
    latents = sched.add_noise(x0b, nzb, timestep.cpu())
    target = sched.training_target(x0b, nzb, timestep)
    with torch.enable_grad():
        pred = tr.model_fn(params, latents, timestep, ctx.cuda().to(BF), True, masks=masks)
        loss = F.mse_loss(pred.float(), target.float()) * sched.training_weight(timestep.cpu())
        loss.backward()
    got = torch.cat([p.grad.float().flatten() for p in params])
    assert rel_l2(got, want) < 1e-2       # p.grad is bf16 (the parameter dtype), `want` fp32
    # one AdamW step moves every B2 entry that has a gradient, and none that was dropped
    tr.optimizer_step(lr=1e-3, weight_decay=0.0)
    ops.sync_check()
    for n in tr.targets:
        moved = (tr.b2[n].float() != b2[n].to(BF).float().cuda())
        assert moved.any() and not moved[masks[n].cuda() == 0].any(), n


def test_real_dims_training_step(env):
    """D = 3072, 24 heads, F = 14336 (TI2V-5B block shapes), 2 layers, S = 320: the production tile paths of every kernel."""
    fg, ops, o, t = env
    ocfg = o.DiTConfig(num_layers=2)
    cfg = fg.WanDiTConfig(num_layers=2)
    w, lora, b2, masks, eng, Stage2Trainer = _setup(fg, o, t, cfg, ocfg)
    shape = (1, 48, 5, 16, 16)
    x0, _, ctx, _ = o.make_inputs(ocfg, shape, text_len=512, live_text=64)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    tr = Stage2Trainer(eng, lora, rank=32)
    tr.load_b2(b2)
    tr.zero_grad()
    loss, pred = tr.step(x0, noise, 500, ctx.cuda(), masks=masks, return_pred=True)
    ops.sync_check()
    loss_ref, pred_ref, grads_ref = _oracle_on_gpu(o, t, ocfg, w, lora, b2, masks, x0, noise, 500, ctx)
    assert rel_l2(pred, pred_ref) < 1e-2
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    worst = max(rel_l2(tr.grad[n], grads_ref[n]) for n in tr.targets)
    print(f"real dims: pred {rel_l2(pred, pred_ref):.3e} loss {float(loss):.6f} vs {float(loss_ref):.6f} worst grad rel-L2 {worst:.3e}")
    assert worst < 5e-2


def test_training_entry_points_reject_bad_arguments(env):
    """Error convention of the C ABI (status code + fgb_last_error -> RuntimeError / ValueError in the shim), no crashes."""
    _, ops, _, _ = env
    q = rnd(100, 128, seed=1)
    out, dq = torch.empty_like(q), torch.empty_like(q)
    lse_bad = torch.zeros(1, 100, dtype=torch.float32, device="cuda")          # stride not a multiple of 64
    with pytest.raises(ValueError):
        ops.attention(q, q, q, out, 1, lse=lse_bad)
    with pytest.raises(ValueError):
        ops.attention_bwd(q, q, q, out, out, lse_bad, dq, dq.clone(), dq.clone(), 1)
    with pytest.raises(ValueError):
        ops.gemm_dgrad(rnd(8, 64), rnd(32, 64), torch.empty(8, 64, dtype=BF, device="cuda"))     # dy cols != w rows
    with pytest.raises(RuntimeError, match="must divide by 64"):
        ops.lora_merge(rnd(100, 256), rnd(32, 256), rnd(100, 32), None, None, torch.empty(100, 256, dtype=BF, device="cuda"))
    with pytest.raises(RuntimeError, match="rank"):
        ops.lora_merge(rnd(128, 256), rnd(24, 256), rnd(128, 24), None, None, torch.empty(128, 256, dtype=BF, device="cuda"))
    with pytest.raises(ValueError):
        ops.lora_wgrad(rnd(64, 128), rnd(32, 32), torch.zeros(128, 32, device="cuda"))            # token counts differ
    with pytest.raises(RuntimeError, match="multiple of 256"):
        ops.ln_bwd(rnd(8, 100), rnd(8, 100), torch.empty(8, 100, dtype=BF, device="cuda"), 1e-6, rnd(100))
    with pytest.raises(ValueError):
        ops.gelu_tanh(rnd(8, 16), torch.empty(8, 8, dtype=BF, device="cuda"))
    ops.sync_check()   # none of the rejected calls launched anything


def test_stage1_training_step_vs_oracle(env):
    """Stage 1 (identity LoRA, training_module.py:200-264): A and B trainable, weight dropout 0.8 on B. Same kernels; dA comes
    from the rank-r factor u = dY·Beff of the dgrad K-extension (fgb_lora_wgrad with transposed output)."""
    fg, ops, o, t = env
    from fairygen_b200 import lora_io
    from fairygen_b200.training import Stage2Trainer
    ocfg = o.TINY
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w, lora = o.make_weights(ocfg, seed=0), o.make_lora(ocfg, rank=32, seed=2)
    masks = t.make_masks_stage1(ocfg)
    eng = fg.WanDiTEngine(cfg, "cuda")
    eng.load_state_dict(w)
    shape = (1, 48, 3, 8, 8)
    x0, _, ctx, _ = o.make_inputs(ocfg, shape, text_len=32, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    tr = Stage2Trainer(eng, lora, rank=32, stage=1)
    assert tr.mask_mul == pytest.approx(5.0)
    tr.zero_grad()
    loss, pred = tr.step(x0, noise, 500, ctx.cuda(), masks=masks, return_pred=True)
    ops.sync_check()
    r = lambda v: v.to(BF).float().cuda()  # noqa: E731
    loss_ref, pred_ref, grads_ref = t.loss_and_grads_stage1({k: r(v) for k, v in w.items()}, ocfg, {k: r(v) for k, v in lora.items()},
                                                            {k: v.cuda() for k, v in masks.items()}, r(x0), r(noise), 500, r(ctx),
                                                            timestep_dtype=BF)
    assert rel_l2(pred, pred_ref) < 1e-2
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    worst_a = max(rel_l2(tr.grad_a[n], grads_ref[f"{n}.lora_A.default.weight"]) for n in tr.targets)
    worst_b = max(rel_l2(tr.grad[n], grads_ref[f"{n}.lora_B.default.weight"]) for n in tr.targets)
    print(f"stage 1: pred {rel_l2(pred, pred_ref):.3e} loss {float(loss):.6f} vs {float(loss_ref):.6f} worst dA {worst_a:.3e} dB {worst_b:.3e}")
    assert worst_a < 5e-2 and worst_b < 5e-2
    for n in tr.targets:
        assert torch.all(tr.grad[n][masks[n].cuda() == 0] == 0), n
    before_a, before_b = tr.a_flat.clone(), tr.b2_flat.clone()
    tr.optimizer_step(lr=1e-3, weight_decay=0.0)
    ops.sync_check()
    assert (tr.a_flat != before_a).any() and (tr.b2_flat != before_b).any()
    sd = lora_io.stage1_state_dict(tr)
    assert set(sd) == set(lora) and all(v.dtype == BF for v in sd.values())
