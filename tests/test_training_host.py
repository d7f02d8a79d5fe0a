"""Host logic of the LoRA trainer (fairygen_b200/training.py) on the CPU: every kernel is replaced by a plain-torch statement of
its contract — backward kernels by torch autograd through the forward contract — and one whole training step (noise / target,
forward with saved activations, loss, hand-ordered backward, LoRA gradients) must reproduce the pinned training oracle, for
stage 2 (lora_B2) and stage 1 (lora_A, lora_B), stored and re-computed activations.  Kernels: tests/test_training_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

from oracle import wan_dit_oracle as o
from oracle import wan_train_oracle as t
from test_engine_host import _emulated_ops as _engine_ops

BF = torch.bfloat16


def _emulated_ops(monkeypatch, holder):
    _engine_ops(monkeypatch)
    from fairygen_b200 import ops
    fwd = {k: getattr(ops, k) for k in ("attention", "rmsnorm_rope", "gemm")}          # the forward contracts just installed

    def ln(x, eps):
        return (x - x.mean(-1, keepdim=True)) * torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)

    def gemm_dgrad(dy, w, dx, u=None, a1=None):
        acc = dy.float() @ w.float()
        if u is not None:
            acc = acc + u.float() @ a1.float()
        dx.copy_(acc.to(BF))
        return dx

    def attention(q, k, v, out, heads, scale=None, lse=None, kmax2=None):
        fwd["attention"](q, k, v, out, heads)
        if lse is not None:                                                              # log2-domain log-sum-exp rows
            qf, kf = (x.float().view(x.shape[0], heads, 128).transpose(0, 1) for x in (q, k))
            lse[:, :q.shape[0]] = torch.logsumexp(qf @ kf.transpose(1, 2) / 128 ** 0.5, -1) * 1.4426950408889634
        return out

    @torch.enable_grad()     # the trainer runs under no_grad; the contract is stated through autograd
    def attention_bwd(q, k, v, out, dout, lse, dq, dk, dv, heads, scale=None, delta=None):
        qs, ks, vs = (x.float().clone().requires_grad_(True) for x in (q, k, v))
        qf, kf, vf = (x.view(x.shape[0], heads, 128).transpose(0, 1) for x in (qs, ks, vs))
        o_ = (torch.softmax(qf @ kf.transpose(1, 2) / 128 ** 0.5, -1) @ vf).transpose(0, 1).reshape(q.shape[0], heads * 128)
        o_.backward(dout.float())
        dq.copy_(qs.grad.to(BF)), dk.copy_(ks.grad.to(BF)), dv.copy_(vs.grad.to(BF))
        return dq, dk, dv

    @torch.enable_grad()     # the trainer runs under no_grad; the contract is stated through autograd
    def ln_bwd(x, dy, out, eps, g0, g1=None, rows_mod0=0, affine=False, dres=None):
        xs = x.float().clone().requires_grad_(True)
        if affine:
            g = g0.float()[None]
        else:
            first = (torch.arange(x.shape[0]) < rows_mod0)[:, None]
            g = 1 + torch.where(first, g0.float()[None], (g0 if g1 is None else g1).float()[None])
        (ln(xs, eps) * g).backward(dy.float())
        out.copy_(((0 if dres is None else dres.float()) + xs.grad).to(BF))
        return out

    @torch.enable_grad()     # the trainer runs under no_grad; the contract is stated through autograd
    def rmsnorm_rope_bwd(x_pre, dy, eps, weight, rope_tab=None, grid=(1, 1, 1), token_offset=0):
        xs = x_pre.float().clone().requires_grad_(True)
        y = xs * torch.rsqrt(xs.pow(2).mean(-1, keepdim=True) + eps) * weight.float()
        if rope_tab is not None:
            f, h, w = grid
            rows, dim = x_pre.shape
            tok = torch.arange(rows) + token_offset
            pos = torch.stack([tok // (h * w), (tok // w) % h, tok % w], 1)
            lanes = torch.tensor([0] * 22 + [1] * 21 + [2] * 21)
            tab = rope_tab[pos[:, lanes], torch.arange(64)]
            z = y.view(rows, dim // 128, 64, 2)
            y = torch.stack([z[..., 0] * tab[:, None, :, 0] - z[..., 1] * tab[:, None, :, 1],
                             z[..., 0] * tab[:, None, :, 1] + z[..., 1] * tab[:, None, :, 0]], -1).reshape(rows, dim)
        y.backward(dy.float())
        dy.copy_(xs.grad.to(BF))
        return dy

    def gelu_tanh(z, h):
        h.copy_(F.gelu(z.float(), approximate="tanh").to(BF))
        return h

    @torch.enable_grad()     # the trainer runs under no_grad; the contract is stated through autograd
    def gelu_tanh_bwd(z, dh, dz):
        zs = z.float().clone().requires_grad_(True)
        F.gelu(zs, approximate="tanh").backward(dh.float())
        dz.copy_(zs.grad.to(BF))
        return dz

    def mul_gate(dx, out, gate0, gate1, rows_gate0):
        first = (torch.arange(dx.shape[0]) < rows_gate0)[:, None]
        out.copy_((dx.float() * torch.where(first, gate0.float()[None], gate1.float()[None])).to(BF))
        return out

    def lora_merge(w, a1, b1, b2, mask, w_eff, mask_mul=2.0, scaling=1.0):
        b = b1.float()
        if b2 is not None:
            b = b + ((b2.float() * (1 if mask is None else mask.float())).to(BF).float() * mask_mul).to(BF).float()
        w_eff.copy_((w.float() + scaling * (b @ a1.float())).to(BF))
        return w_eff

    def lora_b2_eff_batched(b2_flat, mask_flat, table, rank, mask_mul=2.0, scaling=1.0):
        tr = holder["trainer"]                                  # destination pointers -> the trainer's operand buffers
        bufs = [buf for be in tr.b2e for buf in be.values()]
        for off, rows, dst, ld in table.tolist():
            src = b2_flat[off:off + rows * rank].view(rows, rank).float()
            m = 1 if mask_flat is None else mask_flat[off:off + rows * rank].view(rows, rank).float()
            val = (((src * m).to(BF).float() * mask_mul).to(BF).float() * scaling).to(BF)
            buf = next(b for b in bufs if b.data_ptr() <= dst < b.data_ptr() + b.numel() * 2)
            buf.view(-1)[(dst - buf.data_ptr()) // 2:].as_strided((rows, rank), (ld, 1)).copy_(val)

    def lora_wgrad(dy, tt, db, mask=None, mul=1.0, transpose=False):
        g = dy.float().T @ tt.float()                          # [n, r]
        if transpose:
            db += mul * g.T
        else:
            db += mul * g * (1 if mask is None else mask.float())
        return db

    def bernoulli_mask(out_u8, drop_prob, seed):
        out_u8.copy_((torch.rand(out_u8.shape, generator=torch.Generator().manual_seed(seed % (2 ** 31))) > drop_prob).to(torch.uint8))

    def fm_noise_target(x0, noise, sigma, latents, target):
        latents.copy_(((1 - sigma) * x0.float() + sigma * noise.float()).to(BF))
        target.copy_((noise.float() - x0.float()).to(BF))

    def mse_loss_grad(pred, target, weight, loss_f32, dpred=None):
        d = pred.float() - target.float()
        loss_f32.fill_(float(weight * d.pow(2).mean()))
        if dpred is not None:
            dpred.copy_((2 * weight / d.numel() * d).to(BF))

    def unpatchify_bwd(dpred, d_rows, grid):
        C = dpred.shape[0]
        f, h, w = grid
        x = dpred.view(C, f, h, 2, w, 2).permute(1, 2, 4, 3, 5, 0).reshape(f * h * w, 4 * C)      # d_rows[t, y*2C + z*C + c]
        d_rows[:f * h * w, :4 * C] = x
        return d_rows

    def adamw_step(param, grad, m, v, lr, beta1, beta2, eps, weight_decay, step):
        p = param.float() * (1 - lr * weight_decay)
        m.mul_(beta1).add_(grad, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        p = p - lr * (m / (1 - beta1 ** step)) / ((v / (1 - beta2 ** step)).sqrt() + eps)
        param.copy_(p.to(BF))

    for name, fn in list(locals().items()):
        if callable(fn) and name != "ln" and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _trainer(monkeypatch, stage, recompute):
    holder = {}
    _emulated_ops(monkeypatch, holder)
    import fairygen_b200 as fg
    from fairygen_b200 import ops
    from fairygen_b200.training import Stage2Trainer
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w = o.make_weights(o.TINY, seed=0)
    eng = fg.WanDiTEngine.__new__(fg.WanDiTEngine)              # see tests/test_engine_host.py: the constructor refuses the CPU
    eng.cfg, eng.device, eng.ctx, eng.sp = cfg, torch.device("cpu"), None, None
    eng.rope_tab = torch.from_numpy(ops.rope_table(cfg.head_dim))
    eng._init_state()
    eng.load_state_dict(w)
    lora = o.make_lora(o.TINY, rank=32, seed=2)
    tr = Stage2Trainer(eng, lora, rank=32, stage=stage, recompute=recompute)
    holder["trainer"] = tr
    return tr, w, lora


@pytest.mark.parametrize("recompute", [False, True])
def test_stage2_step_matches_the_training_oracle(monkeypatch, recompute):
    tr, w, lora = _trainer(monkeypatch, 2, recompute)
    b2, masks = t.make_b2(o.TINY, rank=32), t.make_masks(o.TINY, rank=32)
    shape = (1, 48, 3, 8, 8)
    x0, _, ctx, _ = o.make_inputs(o.TINY, shape, text_len=32, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    tr.load_b2(b2)
    tr.zero_grad()
    loss, pred = tr.step(x0, noise, 500, ctx, masks=masks, return_pred=True)
    r = lambda v: v.to(BF).float()  # noqa: E731
    loss_ref, pred_ref, grads_ref = t.loss_and_grads({k: r(v) for k, v in w.items()}, o.TINY, {k: r(v) for k, v in lora.items()},
                                                     {k: r(v) for k, v in b2.items()}, masks, r(x0), r(noise), 500, r(ctx), timestep_dtype=BF)
    assert rel(pred.float(), pred_ref) < 1e-2
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    for name in tr.targets:
        assert rel(tr.grad[name], grads_ref[name]) < 5e-2, (name, rel(tr.grad[name], grads_ref[name]))
        assert torch.all(tr.grad[name][masks[name] == 0] == 0), name
    before = tr.b2_flat.clone()
    tr.optimizer_step(lr=1e-3, weight_decay=0.0)
    assert (tr.b2_flat != before).any()


def test_stage1_step_matches_the_training_oracle(monkeypatch):
    tr, w, lora = _trainer(monkeypatch, 1, False)
    masks = t.make_masks_stage1(o.TINY, rank=32)
    shape = (1, 48, 2, 6, 10)
    x0, _, ctx, _ = o.make_inputs(o.TINY, shape, text_len=24, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    tr.zero_grad()
    loss, pred = tr.step(x0, noise, 37, ctx, masks=masks, return_pred=True)
    r = lambda v: v.to(BF).float()  # noqa: E731
    loss_ref, pred_ref, grads_ref = t.loss_and_grads_stage1({k: r(v) for k, v in w.items()}, o.TINY, {k: r(v) for k, v in lora.items()}, masks,
                                                            r(x0), r(noise), 37, r(ctx), timestep_dtype=BF)
    assert rel(pred.float(), pred_ref) < 1e-2
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    worst_a = max(rel(tr.grad_a[n], grads_ref[n + ".lora_A.default.weight"]) for n in tr.targets)
    worst_b = max(rel(tr.grad[n], grads_ref[n + ".lora_B.default.weight"]) for n in tr.targets)
    assert worst_a < 5e-2 and worst_b < 5e-2, (worst_a, worst_b)
