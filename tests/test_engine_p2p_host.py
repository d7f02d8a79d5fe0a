"""Multi-process CPU test (2 and 4 ranks) of the WHOLE DiT forward under the PEER-STORE Ulysses exchange (the default on GPUs, USP:125-146 as our
kernels do it): the product's engine, `PeerArena` (offsets of recv | o | stats | flags | status, handle exchange over
torch.distributed), `SequenceParallel.attention` (fused send and the scatter-kernel variant), the epoch protocol of the two
barriers and `SequenceParallel.check()` run for real between the processes.  What stands in for the GPU:

* the arena's backing store is a file mapped into every process (`PeerArena._allocate` hook; fgb_ipc_export / fgb_ipc_open become
  "name the file" / "map it"), so a peer pointer + offset is real shared memory;
* every kernel is a plain-torch statement of its contract in include/fairygen_b200.h (tests/test_engine_host.py for the local
  ones; the exchange kernels below: fgb_gemm_qkv_scatter, fgb_sp_stats_barrier, fgb_recv_norm_rope, fgb_attn_fwd_scatter,
  fgb_sp_barrier[_status], fgb_rmsnorm_rope_scatter, fgb_sp_scatter_heads) that reads and writes THROUGH the pointers it is given.

The `-m gpu` twins (tests/test_sp_gpu.py) need two GPUs and are skipped on a one-GPU box; this one runs everywhere."""
import math
import os
import socket
import sys
import time

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BF = torch.bfloat16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Patch:
    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


class _SharedArenas:
    """File-backed stand-in for CUDA IPC: every arena is a file in `directory`, mapped shared; pointers are the real addresses of
    the mappings in THIS process, so the product's pointer arithmetic (base + offset) is exercised as is."""

    def __init__(self, directory: str, rank: int):
        self.dir, self.rank, self.generation, self.maps = directory, rank, 0, []

    def _map(self, name: str, nbytes: int) -> torch.Tensor:
        path = os.path.join(self.dir, name)
        if not os.path.exists(path):
            with open(path, "wb") as f:
                f.truncate(nbytes)                      # zero-filled, like the torch.zeros of the product
        t = torch.from_file(path, shared=True, size=nbytes, dtype=torch.uint8)
        self.maps.append((name, t))
        return t

    def allocate(self, nbytes: int, device) -> torch.Tensor:
        self.generation += 1
        return self._map(f"arena_r{self.rank}_g{self.generation}", nbytes)

    def export(self, t: torch.Tensor):
        for name, m in self.maps:
            if m.data_ptr() == t.data_ptr():
                return f"{name}:{m.numel()}".encode().ljust(64, b"\0"), 0
        raise AssertionError("ipc_export of a tensor that is not an arena")

    def open(self, device, handle: bytes, offset: int) -> int:
        name, nbytes = handle.rstrip(b"\0").decode().split(":")
        return self._map(name, int(nbytes)).data_ptr() + offset

    def view(self, ptr: int, shape, dtype) -> torch.Tensor:
        nbytes = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        for _, m in self.maps:
            off = ptr - m.data_ptr()
            if 0 <= off and off + nbytes <= m.numel():
                return m[off:off + nbytes].view(dtype).view(*shape)
        raise AssertionError(f"pointer {ptr:#x} (+{nbytes} bytes) is outside every mapped arena")


def _rope(y, rope_tab, grid, token0=0):
    """3-D RoPE of DIT:91-96 on y [rows, heads*128] fp32 for tokens token0 .. ; rows past the grid are left alone (padding)."""
    f, h, w = grid
    rows, dim = y.shape
    t = torch.arange(rows) + token0
    live = t < f * h * w
    t = torch.where(live, t, torch.zeros_like(t))
    pos = torch.stack([t // (h * w), (t // w) % h, t % w], 1)
    lanes = torch.tensor([0] * 22 + [1] * 21 + [2] * 21)
    tab = rope_tab[pos[:, lanes], torch.arange(64)]                          # [rows, 64, (cos, sin)]
    z = y.view(rows, dim // 128, 64, 2)
    re = z[..., 0] * tab[:, None, :, 0] - z[..., 1] * tab[:, None, :, 1]
    im = z[..., 0] * tab[:, None, :, 1] + z[..., 1] * tab[:, None, :, 0]
    return torch.where(live[:, None], torch.stack([re, im], -1).reshape(rows, dim), y)


def _exchange_kernels(ops, shm):
    """Contract statements of the exchange kernels, working through the peer pointers."""

    def epoch_exchange(flag_ptrs, world, rank, epoch, status, limit_s=float(os.environ.get("FGB_TEST_BARRIER_LIMIT_S", "60"))):
        for q in range(world):                                               # publish: peer q's slot [rank]
            shm.view(flag_ptrs[q], (64,), torch.int32)[rank] = epoch
        mine = shm.view(flag_ptrs[rank], (64,), torch.int32)
        t0 = time.time()
        for q in range(world):
            while int(mine[q]) - epoch < 0:
                if status is not None and int(status[0]) != 0:
                    return
                if time.time() - t0 > limit_s:
                    assert status is not None, "barrier timed out without a status word (the kernel would trap)"
                    status[0] = epoch
                    return
                time.sleep(0.0002)

    def sp_barrier(device, flag_ptrs, world, rank, epoch, status=None, timeout_clocks=0):
        if timeout_clocks == 0:
            epoch_exchange(flag_ptrs, world, rank, epoch, status)
        else:
            epoch_exchange(flag_ptrs, world, rank, epoch, status, timeout_clocks / 1.5e9)

    def gemm_qkv_scatter(a, w, bias, dim, peer_recv_ptrs, world, rank, rowsq, sk_ws=None):
        m = a.shape[0]
        y = (a.float() @ w.float().T + bias.float()).to(BF)                  # [m, q|k|v]
        wloc = dim // world
        for q in range(world):
            recv = shm.view(peer_recv_ptrs[q], (world * m, 3 * wloc), BF)
            for g in range(3):
                recv[rank * m:(rank + 1) * m, g * wloc:(g + 1) * wloc] = y[:, g * dim + q * wloc:g * dim + (q + 1) * wloc]
        rowsq.view(2, m)[0] += y[:, :dim].float().pow(2).sum(-1)
        rowsq.view(2, m)[1] += y[:, dim:2 * dim].float().pow(2).sum(-1)

    def sp_stats_barrier(device, flag_ptrs, stats_ptrs, rowsq, rows, s_pad, kmax2, hpr, world, rank, epoch, status=None):
        for q in range(world):
            shm.view(stats_ptrs[q], (2, s_pad), torch.float32)[:, rank * rows:(rank + 1) * rows] = rowsq.view(2, rows)
        rowsq.zero_()
        kmax2[:hpr].zero_()
        epoch_exchange(flag_ptrs, world, rank, epoch, status)

    def recv_norm_rope(recv, tokens, hpr, stats, dim, eps, wq, wk, rope_tab, grid, kmax2, qmax2=None):
        wloc = hpr * 128
        for g, weight in ((0, wq), (1, wk)):
            x = recv[:, g * wloc:(g + 1) * wloc].float()
            rs = torch.rsqrt(stats.view(2, -1)[g] / dim + eps)[:, None]      # the statistics are those of the FULL row
            y = ((x * rs).to(BF).float() * weight.float()).to(BF).float()
            recv[:, g * wloc:(g + 1) * wloc] = _rope(y, rope_tab, grid).to(BF)
        kmax2.copy_(recv[:tokens, wloc:2 * wloc].float().view(tokens, hpr, 128).pow(2).sum(-1).max(0).values)
        if qmax2 is not None:
            qmax2.copy_(recv[:, :wloc].float().view(recv.shape[0], hpr, 128).pow(2).sum(-1).max(0).values)

    def attention_scatter(q, k, v, o_peer_ptrs, ldo, rows_per_peer, col_offset, heads, scale=None, kmax2=None, lse=None, qmax2=None):
        qf, kf, vf = (t.float().view(t.shape[0], heads, 128).transpose(0, 1) for t in (q, k, v))
        p = torch.softmax(qf @ kf.transpose(1, 2) / 128 ** 0.5, -1)
        out = (p @ vf).transpose(0, 1).reshape(q.shape[0], heads * 128).to(BF)
        if lse is not None:                                                   # log2-domain log-sum-exp rows of the local heads
            lse[:, :q.shape[0]] = torch.logsumexp(qf @ kf.transpose(1, 2) / 128 ** 0.5, -1) * 1.4426950408889634
        for peer, ptr in enumerate(o_peer_ptrs):                              # row of global token t -> its owner, at the head columns
            o = shm.view(ptr, (rows_per_peer, ldo), BF)
            o[:, col_offset:col_offset + heads * 128] = out[peer * rows_per_peer:(peer + 1) * rows_per_peer]

    def sp_scatter_heads(x, peer_ptrs, heads, groups, world, rank, group_first=0, groups_total=None):
        groups_total = groups if groups_total is None else groups_total
        s_local, hpr = x.shape[0], heads // world
        xv = x.reshape(s_local, groups, world, hpr * 128)
        for q in range(world):
            recv = shm.view(peer_ptrs[q], (world * s_local, groups_total, hpr * 128), BF)
            recv[rank * s_local:(rank + 1) * s_local, group_first:group_first + groups] = xv[:, :, q]

    def sp_return_heads(x, peer_ptrs, ld_dst, rows, heads, groups, world, rank):
        hpr = heads // world                                                  # x: this rank's heads, ALL tokens -> the token owners
        xv = x.reshape(x.shape[0], groups, hpr * 128)
        for peer, ptr in enumerate(peer_ptrs):
            dst = shm.view(ptr, (rows, ld_dst), BF)[:, :groups * heads * 128].view(rows, groups, heads * 128)
            dst[:, :, rank * hpr * 128:(rank + 1) * hpr * 128] = xv[peer * rows:(peer + 1) * rows]

    def rmsnorm_rope_scatter(x, eps, weight, rope_tab, grid, token_offset, peer_ptrs, world, rank, group, groups_total):
        xf = x.float()
        y = ((xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(BF).float() * weight.float()).to(BF).float()
        y = _rope(y, rope_tab, grid, token_offset).to(BF)
        sp_scatter_heads(y, peer_ptrs, x.shape[1] // 128, 1, world, rank, group, groups_total)

    for name, fn in list(locals().items()):
        if callable(fn) and hasattr(ops, name):
            setattr(ops, name, fn)
    ops.ipc_export = shm.export
    ops.ipc_open = shm.open
    ops.ipc_close = lambda device, ptr, off: None


def _worker(rank, world, port, dims, shape, fused, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), FGB_SP_FUSED="1" if fused else "0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import fairygen_b200 as fg
    from fairygen_b200 import ops, sp as spmod
    from oracle import wan_dit_oracle as o
    from test_engine_host import _emulated_ops
    from test_engine_sp_gloo import _bare_engine

    _emulated_ops(_Patch)
    shm = _SharedArenas(out_dir, rank)
    _exchange_kernels(ops, shm)
    spmod.PeerArena._allocate = staticmethod(shm.allocate)

    dim, ffn, heads, text = dims
    ocfg = o.DiTConfig(dim=dim, ffn_dim=ffn, text_dim=text, num_heads=heads, num_layers=2)
    cfg = fg.WanDiTConfig(dim=dim, ffn_dim=ffn, text_dim=text, num_heads=heads, num_layers=2)
    w = o.make_weights(ocfg, seed=0)
    lat, _, cp, _ = o.make_inputs(ocfg, shape, text_len=32, live_text=8)
    single = _bare_engine(fg, ops, cfg, None)
    single.load_state_dict(w)
    par = _bare_engine(fg, ops, cfg, fg.SequenceParallel(exchange="p2p"))
    assert par.sp.fused_send == fused
    par.load_state_dict(w)
    res = {"shape": None, "vs_single": [], "vs_oracle": []}
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())  # noqa: E731
    for ts_val, per_token in ((900.0, True), (37.0, False)):                 # two forwards: the epochs keep counting, the arena is reused
        ts = torch.tensor([ts_val])
        ref = single.forward(lat.to(BF), ts, cp.to(BF), per_token)
        out = par.forward(lat.to(BF), ts, cp.to(BF), per_token)
        with torch.no_grad():
            want = o.dit_forward({k: v.to(BF).float() for k, v in w.items()}, ocfg, lat.to(BF).float(), ts, cp.to(BF).float(), per_token)
        res["shape"] = tuple(out.shape)
        res["vs_single"].append(rel(out.float(), ref.float()))
        res["vs_oracle"].append(rel(out.float(), want))
    par.sp.check()                                                           # no barrier gave up
    ar = par.sp.arena
    res.update(epoch=ar.epoch, status=int(ar.status.item()),
               flags=[shm.view(ar.flag_ptrs[which][rank], (64,), torch.int32)[:world].tolist() for which in (0, 1)],
               untouched=int(shm.view(ar.flag_ptrs[0][rank], (64,), torch.int32)[world:].abs().sum()
                             + shm.view(ar.flag_ptrs[1][rank], (64,), torch.int32)[world:].abs().sum()))
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dims,shape,fused", [
    (2, (256, 512, 2, 128), (1, 48, 4, 8, 8), True),        # 1 head per rank, S = 64 (even split); GEMM-with-send + stats barrier
    (2, (256, 512, 2, 128), (1, 48, 4, 8, 8), False),       # the same through the norm-and-send / scatter kernels
    (2, (512, 512, 4, 128), (1, 48, 3, 10, 14), True),      # 2 heads per rank, S = 105 (53 + 52 and a pad row)
    (2, (512, 512, 4, 128), (1, 48, 3, 10, 14), False),
    (4, (512, 512, 4, 128), (1, 48, 3, 10, 14), True),      # 4 ranks: 27 rows each, 3 pad rows on the last
])
def test_peer_store_exchange_forward(tmp_path, world, dims, shape, fused):
    mp.spawn(_worker, args=(world, _free_port(), dims, shape, fused, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["shape"] == shape
        assert max(res["vs_single"]) < 4e-3, res          # same contracts; only the summation order inside attention differs
        assert max(res["vs_oracle"]) < 1e-2, res          # north star: rel L2 <= 1e-2 per forward
        # one epoch per block and forward (2 layers x 2 forwards), both flag sets at it on every rank, nothing else written
        assert res["epoch"] == 4 and res["status"] == 0 and res["flags"] == [[4] * world, [4] * world] and res["untouched"] == 0, res


def _train_worker(rank, world, port, shape, recompute, out_dir):
    """One stage-2 LoRA training step (forward with saved activations, loss, hand-ordered backward) with the tokens of the video
    split over the ranks: the exchange runs forward (q|k|v out, O back) AND backward (O, dO out, dq|dk|dv back through
    fgb_sp_return_heads into the arena's gradient region), the B2 gradients are summed over the group."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import fairygen_b200 as fg
    from fairygen_b200 import ops, sp as spmod
    from fairygen_b200.training import Stage2Trainer
    from oracle import wan_dit_oracle as o
    from oracle import wan_train_oracle as t
    from test_engine_sp_gloo import _bare_engine
    from test_training_host import _emulated_ops as _training_ops

    holder = {}
    _training_ops(_Patch, holder)
    shm = _SharedArenas(out_dir, rank)
    _exchange_kernels(ops, shm)
    spmod.PeerArena._allocate = staticmethod(shm.allocate)

    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w, lora = o.make_weights(o.TINY, seed=0), o.make_lora(o.TINY, rank=32, seed=2)
    b2, masks = t.make_b2(o.TINY, rank=32), t.make_masks(o.TINY, rank=32)
    x0, _, ctx, _ = o.make_inputs(o.TINY, shape, text_len=32, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))

    def run(sp):
        eng = _bare_engine(fg, ops, cfg, sp)
        eng.load_state_dict(w)
        tr = Stage2Trainer(eng, lora, rank=32, stage=2, recompute=recompute)
        holder["trainer"] = tr
        tr.load_b2(b2)
        tr.zero_grad()
        loss, pred = tr.step(x0, noise, 500, ctx, masks=masks, return_pred=True)
        return tr, float(loss), pred.float().clone(), {n: tr.grad[n].clone() for n in tr.targets}

    _, loss1, pred1, grads1 = run(None)
    par = fg.SequenceParallel(exchange="p2p")
    trp, lossp, predp, gradsp = run(par)
    par.check()
    r = lambda v: v.to(BF).float()  # noqa: E731
    _, pred_ref, grads_ref = t.loss_and_grads({k: r(v) for k, v in w.items()}, o.TINY, {k: r(v) for k, v in lora.items()},
                                              {k: r(v) for k, v in b2.items()}, masks, r(x0), r(noise), 500, r(ctx), timestep_dtype=BF)
    rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))  # noqa: E731
    torch.save({"loss": (lossp, loss1), "pred_vs_single": rel(predp, pred1), "pred_vs_oracle": rel(predp, pred_ref),
                "grad_vs_single": max(rel(gradsp[n], grads1[n]) for n in trp.targets),
                "grad_vs_oracle": max(rel(gradsp[n], grads_ref[n]) for n in trp.targets),
                "masked_zero": all(bool(torch.all(gradsp[n][masks[n] == 0] == 0)) for n in trp.targets),
                "epoch": par.arena.epoch, "status": int(par.arena.status.item()),
                "flags": [shm.view(par.arena.flag_ptrs[which][rank], (64,), torch.int32)[:world].tolist() for which in (0, 1)]},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,recompute", [((1, 48, 3, 10, 14), False),      # S = 105: ragged split, one padded row
                                             ((1, 48, 4, 8, 8), True)])       # S = 64, activations re-computed per block (PIPE:1348-1360)
def test_peer_store_exchange_training_step_world2(tmp_path, shape, recompute):
    world = 2
    mp.spawn(_train_worker, args=(world, _free_port(), shape, recompute, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["pred_vs_single"] < 4e-3 and res["pred_vs_oracle"] < 1e-2, res
        assert abs(res["loss"][0] - res["loss"][1]) < 2e-3 * abs(res["loss"][1]), res
        assert res["grad_vs_single"] < 3e-2 and res["grad_vs_oracle"] < 5e-2 and res["masked_zero"], res
        # per block one exchange forward and one backward (+ one more forward when the block is re-computed)
        epochs = 2 * (3 if recompute else 2)
        assert res["epoch"] == epochs and res["status"] == 0 and res["flags"] == [[epochs] * world, [epochs] * world], res


def _layout_worker(rank, world, port, out_dir):
    """The headline layout of the 8-GPU line in small: CFG pair x Ulysses groups with the PEER-STORE exchange (two disjoint SP
    groups, each with its own arenas and epochs; the pair exchanges predictions) — a 2-step CFG denoise must equal the sequential
    single-process loop (PIPE:285-309)."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import fairygen_b200 as fg
    from fairygen_b200 import ops, scheduler, sp as spmod
    from fairygen_b200.cfg_parallel import Layout, ParallelContext, denoise_shots
    from oracle import wan_dit_oracle as o
    from test_engine_host import _emulated_ops
    from test_engine_sp_gloo import _bare_engine

    _emulated_ops(_Patch)
    shm = _SharedArenas(out_dir, rank)
    _exchange_kernels(ops, shm)
    spmod.PeerArena._allocate = staticmethod(shm.allocate)

    def step_fused(self, latents, noise_pos, noise_neg, cfg_scale, index, first_frame_latents=None, to_final=False):
        ops.cfg_fm_step(latents, noise_pos, noise_neg, first_frame_latents, float(cfg_scale), self.sigma_delta(index, to_final))
        return latents

    scheduler.FlowMatchScheduler.step_fused = step_fused          # without the CUDA-only guard (tests/test_host_logic.py checks it)
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w = o.make_weights(o.TINY, seed=0)
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 10, 14), text_len=24, live_text=8)       # S = 105: ragged over 2 ranks
    shot = dict(latents=lat, context_pos=cp, context_neg=cn, first_frame_latents=z0)
    ctx = ParallelContext(Layout(world, 1, 2, 2), exchange="p2p")
    par = ctx.sequence_parallel()
    assert par.exchange == "p2p" and par.world == 2
    eng = _bare_engine(fg, ops, cfg, par)
    eng.load_state_dict(w)
    done = denoise_shots(ctx, lambda: fg.WanDenoiser(eng, 2, cfg_scale=5.0, sigma_shift=5.0), [shot])
    single = _bare_engine(fg, ops, cfg, None)
    single.load_state_dict(w)
    ref = fg.WanDenoiser(single, 2, cfg_scale=5.0, sigma_shift=5.0)(lat, cp, cn, z0)
    errs = [float((out.double() - ref.double()).norm() / ref.double().norm()) for _, out in done]
    torch.save({"errs": errs, "epoch": par.arena.epoch, "status": int(par.arena.status.item()), "coords": ctx.layout.coords(rank)},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_cfg_pair_times_peer_store_groups_world4(tmp_path):
    world = 4
    mp.spawn(_layout_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["errs"] and all(e < 5e-3 for e in res["errs"]), res
        assert res["epoch"] == 2 * 2 and res["status"] == 0, res        # each rank runs ONE forward per step: 2 steps x 2 blocks


def _lost_peer_worker(rank, world, port, out_dir):
    """Rank 1 maps its arena and then never launches the forward (a dead rank): rank 0's first barrier gives up, writes its epoch
    to the status word, every later barrier of the forward returns at once, and the host check names rank and epoch."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), FGB_TEST_BARRIER_LIMIT_S="1.0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import fairygen_b200 as fg
    from fairygen_b200 import ops, sp as spmod
    from oracle import wan_dit_oracle as o
    from test_engine_host import _emulated_ops
    from test_engine_sp_gloo import _bare_engine

    _emulated_ops(_Patch)
    shm = _SharedArenas(out_dir, rank)
    _exchange_kernels(ops, shm)
    spmod.PeerArena._allocate = staticmethod(shm.allocate)
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    shape = (1, 48, 4, 8, 8)
    lat, z0, cp, cn = o.make_inputs(o.TINY, shape, text_len=32, live_text=8)
    par = fg.SequenceParallel(exchange="p2p")
    eng = _bare_engine(fg, ops, cfg, par)
    eng.load_state_dict(o.make_weights(o.TINY, seed=0))
    res = {}
    waits = []
    for name in ("sp_barrier", "sp_stats_barrier"):
        def timed(*args, _fn=getattr(ops, name), **kwargs):
            t0 = time.time()
            _fn(*args, **kwargs)
            waits.append(time.time() - t0)
        setattr(ops, name, timed)
    if rank == 0:
        out = eng.forward(lat.to(BF), torch.tensor([900.0]), cp.to(BF), True)
        res["waits"] = waits
        res["shape"] = tuple(out.shape)
        try:
            par.check()
            res["raised"] = None
        except RuntimeError as e:
            res["raised"] = str(e)
        den = fg.WanDenoiser.__new__(fg.WanDenoiser)          # the denoiser's end-of-video check sees the same word
        den.engine, den._host_contexts = eng, []
        den.scheduler = type("S", (), {"timesteps": []})()
        try:
            den(lat, cp, None, steps=range(0))
            res["denoiser_raised"] = False
        except RuntimeError:
            res["denoiser_raised"] = True
    else:
        par.peer_arena(32, cfg.num_heads, torch.device("cpu"))               # maps its arena (collective), then goes silent ...
        rows = torch.zeros(32, cfg.out_dim * 4, dtype=BF)
        par.all_gather_rows(rows, torch.empty(64, cfg.out_dim * 4, dtype=BF))   # ... except for the library collective at the end
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_a_lost_peer_is_reported_not_waited_for(tmp_path):
    world = 2
    mp.spawn(_lost_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = torch.load(os.path.join(tmp_path, "r0.pt"))
    assert res["shape"] == (1, 48, 4, 8, 8)
    waits = res["waits"]                                  # 2 blocks x 2 barriers: ONE waited out its limit (1 s here), the others did not
    assert len(waits) == 4 and waits[0] >= 0.9 and all(w < 0.5 for w in waits[1:]), waits
    assert res["raised"] is not None and "rank 0 of 2" in res["raised"] and "epoch 1 " in res["raised"], res
    assert res["denoiser_raised"], res
