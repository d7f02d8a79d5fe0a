"""Race check of the peer-store Ulysses exchange at the protocol level (compute-sanitizer's racecheck is closed on this pool and
cannot see across GPUs anyway).  The forward of tests/test_engine_p2p_host.py runs between CPU processes with a recorder on: every
kernel launch of the engine that touches arena memory (its own `recv`, `o`, `stats`) and every REMOTE store of the exchange kernels
(into the peers' `recv` rows, `o` columns, `stats` windows) is logged in launch order together with the two barriers of each block.
The checker then forgets the actual timing and asks the only question that matters on the GPUs: is every pair of overlapping
accesses from different ranks ordered by the barriers (vector clocks: a barrier joins the clocks all ranks had when they
published its epoch)?  A launch order in sp.py / engine.py that lets a peer's next send overwrite rows still being read, or an
attention epilogue store into an `o` that its owner has not consumed, fails here on any machine — and so does a log with a barrier
taken out, which is how the checker itself is tested."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_engine_p2p_host import BF, REPO, _exchange_kernels, _free_port, _Patch, _SharedArenas


class _Recorder:
    """engine.timer stand-in: sees every named launch of the forward (engine._k) with its arguments."""

    def __init__(self, log, arena_of):
        self.log, self.arena_of = log, arena_of

    def call(self, name, fn, *args, **kwargs):
        ar = self.arena_of()
        if ar is not None:
            rects = [r for t in list(args) + list(kwargs.values()) if torch.is_tensor(t) for r in _rects(ar, t)]
            if rects:
                self.log.append(("acc", name, rects))
        return fn(*args, **kwargs)


def _regions(ar):
    s_pad, wloc3 = ar.recv.shape
    regions = {"recv": (ar.off_recv, wloc3 * 2, s_pad), "o": (ar.off_o, ar.o.shape[1] * 2, ar.o.shape[0]),
               "stats": (ar.off_stats, s_pad * 4, 2)}
    if ar.dqkv is not None:                                                  # training: the gradient matrix the peers return into
        regions["dqkv"] = (ar.off_dqkv, ar.dqkv.shape[1] * 2, ar.dqkv.shape[0])
    return regions


def _rects(ar, t):
    """The part of THIS rank's arena a local tensor argument covers, as (owner, region, row0, row1, byte0, byte1)."""
    off = t.data_ptr() - ar.buf.data_ptr()
    if not (0 <= off < ar.buf.numel()) or t.numel() == 0:
        return []
    for name, (base, pitch, rows) in _regions(ar).items():
        if base <= off < base + pitch * rows:
            rel = off - base
            if t.dim() == 2 and t.stride(0) * t.element_size() == pitch:
                return [(ar.rank, name, rel // pitch, rel // pitch + t.shape[0], rel % pitch, rel % pitch + t.shape[1] * t.element_size())]
            return [(ar.rank, name, 0, rows, 0, pitch)]                      # anything else: the whole region (conservative)
    return []


def _recording_exchange(ops, shm, log, par):
    """Wrap the exchange kernels of test_engine_p2p_host: same contracts, plus a log of their remote stores and barriers."""
    _exchange_kernels(ops, shm)
    inner = {n: getattr(ops, n) for n in ("gemm_qkv_scatter", "sp_stats_barrier", "attention_scatter", "sp_barrier", "sp_scatter_heads",
                                          "rmsnorm_rope_scatter")}
    which_set = lambda flag_ptrs: [fp[0] for fp in par.arena.flag_ptrs].index(flag_ptrs[0])  # noqa: E731

    def gemm_qkv_scatter(a, w, bias, dim, peer_recv_ptrs, world, rank, rowsq, sk_ws=None):
        m, pitch = a.shape[0], 3 * (dim // world) * 2
        log.append(("acc", "gemm_qkv_scatter", [(q, "recv", rank * m, (rank + 1) * m, 0, pitch) for q in range(world)]))
        return inner["gemm_qkv_scatter"](a, w, bias, dim, peer_recv_ptrs, world, rank, rowsq, sk_ws)

    def sp_scatter_heads(x, peer_ptrs, heads, groups, world, rank, group_first=0, groups_total=None):
        total = groups if groups_total is None else groups_total
        m, gbytes = x.shape[0], (heads // world) * 128 * 2
        log.append(("acc", "sp_scatter_heads", [(q, "recv", rank * m, (rank + 1) * m, group_first * gbytes, (group_first + groups) * gbytes)
                                                for q in range(world)]))
        assert total * gbytes == par.arena.recv.shape[1] * 2
        return inner["sp_scatter_heads"](x, peer_ptrs, heads, groups, world, rank, group_first, groups_total)

    def sp_stats_barrier(device, flag_ptrs, stats_ptrs, rowsq, rows, s_pad, kmax2, hpr, world, rank, epoch, status=None):
        log.append(("acc", "sp_stats_barrier", [(q, "stats", 0, 2, rank * rows * 4, (rank + 1) * rows * 4) for q in range(world)]))
        log.append(("bar", which_set(flag_ptrs), epoch))                     # the push precedes the publish inside the kernel
        return inner["sp_stats_barrier"](device, flag_ptrs, stats_ptrs, rowsq, rows, s_pad, kmax2, hpr, world, rank, epoch, status)

    def attention_scatter(q, k, v, o_peer_ptrs, ldo, rows_per_peer, col_offset, heads, **kw):
        log.append(("acc", "attention_scatter", [(p, "o", 0, rows_per_peer, col_offset * 2, (col_offset + heads * 128) * 2)
                                                 for p in range(len(o_peer_ptrs))]))
        return inner["attention_scatter"](q, k, v, o_peer_ptrs, ldo, rows_per_peer, col_offset, heads, **kw)

    def sp_barrier(device, flag_ptrs, world, rank, epoch, status=None, timeout_clocks=0):
        log.append(("bar", which_set(flag_ptrs), epoch))
        return inner["sp_barrier"](device, flag_ptrs, world, rank, epoch, status, timeout_clocks)

    inner["sp_return_heads"] = ops.sp_return_heads

    def sp_return_heads(x, peer_ptrs, ld_dst, rows, heads, groups, world, rank):
        hb = (heads // world) * 128 * 2
        log.append(("acc", "sp_return_heads", [(p, "dqkv", 0, rows, (g * heads * 128) * 2 + rank * hb, (g * heads * 128) * 2 + (rank + 1) * hb)
                                               for p in range(world) for g in range(groups)]))
        return inner["sp_return_heads"](x, peer_ptrs, ld_dst, rows, heads, groups, world, rank)

    for fn in (gemm_qkv_scatter, sp_scatter_heads, sp_stats_barrier, attention_scatter, sp_barrier, sp_return_heads):
        setattr(ops, fn.__name__, fn)
    # fgb_rmsnorm_rope_scatter sends through the same layout as fgb_sp_scatter_heads: route it through the logging version
    # (its contract statement calls sp_scatter_heads by closure, so re-state the send here)
    def rmsnorm_rope_scatter(x, eps, weight, rope_tab, grid, token_offset, peer_ptrs, world, rank, group, groups_total):
        m, gbytes = x.shape[0], (x.shape[1] // 128 // world) * 128 * 2
        log.append(("acc", "rmsnorm_rope_scatter", [(q, "recv", rank * m, (rank + 1) * m, group * gbytes, (group + 1) * gbytes)
                                                    for q in range(world)]))
        return inner["rmsnorm_rope_scatter"](x, eps, weight, rope_tab, grid, token_offset, peer_ptrs, world, rank, group, groups_total)

    ops.rmsnorm_rope_scatter = rmsnorm_rope_scatter


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
    spmod.PeerArena._allocate = staticmethod(shm.allocate)
    par = fg.SequenceParallel(exchange="p2p")
    log = []
    _recording_exchange(ops, shm, log, par)
    dim, ffn, heads, text = dims
    ocfg = o.DiTConfig(dim=dim, ffn_dim=ffn, text_dim=text, num_heads=heads, num_layers=3)
    cfg = fg.WanDiTConfig(dim=dim, ffn_dim=ffn, text_dim=text, num_heads=heads, num_layers=3)
    eng = _bare_engine(fg, ops, cfg, par)
    eng.load_state_dict(o.make_weights(ocfg, seed=0))
    eng.timer = _Recorder(log, lambda: par.arena)
    lat, _, cp, cn = o.make_inputs(ocfg, shape, text_len=32, live_text=8)
    for ctx in (cp, cn):                                                     # two forwards back to back, as a CFG step launches them
        eng.forward(lat.to(BF), torch.tensor([900.0]), ctx.to(BF), True)
    par.check()
    torch.save(log, os.path.join(out_dir, f"log{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def hazards(logs):
    """logs[rank] = launch-ordered events of that rank.  Returns the unordered conflicting pairs (empty = race-free)."""
    world = len(logs)
    bars = [[e for e in log if e[0] == "bar"] for log in logs]
    assert all(b == bars[0] for b in bars), "ranks disagree on the barrier sequence (a deadlock on the GPUs)"
    clocks = [[0] * world for _ in range(world)]
    stamped = []                                                             # (rank, clock snapshot, name, rects)
    cursor = [0] * world
    for _ in range(len(bars[0]) + 1):
        published = []
        for r in range(world):                                               # run every rank up to its next barrier
            while cursor[r] < len(logs[r]) and logs[r][cursor[r]][0] != "bar":
                _, name, rects = logs[r][cursor[r]]
                clocks[r][r] += 1
                stamped.append((r, list(clocks[r]), name, rects))
                cursor[r] += 1
            clocks[r][r] += 1
            published.append(list(clocks[r]))                                # the clock this rank publishes its epoch with
            cursor[r] += 1
        joined = [max(p[i] for p in published) for i in range(world)]       # leaving the barrier: everyone has seen everyone's publish
        for r in range(world):
            clocks[r] = [max(a, b) for a, b in zip(clocks[r], joined)]
    before = lambda x, y: x[1][x[0]] <= y[1][x[0]]                           # noqa: E731   x happened before y
    bad = []
    for i, x in enumerate(stamped):
        for y in stamped[i + 1:]:
            if x[0] == y[0] or before(x, y) or before(y, x):                # same rank = same stream = launch order
                continue
            for (o1, g1, r0, r1, c0, c1) in x[3]:
                for (o2, g2, s0, s1, d0, d1) in y[3]:
                    if o1 == o2 and g1 == g2 and r0 < s1 and s0 < r1 and c0 < d1 and d0 < c1:
                        bad.append((f"rank {x[0]} {x[2]}", f"rank {y[0]} {y[2]}", f"arena of rank {o1}: {g1} rows [{max(r0, s0)}, {min(r1, s1)})"))
    return bad


def _without_barrier(logs, which, nth):
    """The same logs with the nth barrier on flag set `which` taken out on every rank."""
    out = []
    for log in logs:
        seen, kept = 0, []
        for e in log:
            if e[0] == "bar" and e[1] == which:
                seen += 1
                if seen == nth:
                    continue
            kept.append(e)
        out.append(kept)
    return out


@pytest.mark.parametrize("world,dims,shape,fused", [
    (2, (256, 512, 2, 128), (1, 48, 3, 10, 14), True),       # GEMM-with-send, statistics with barrier 0, receiver norm
    (2, (256, 512, 2, 128), (1, 48, 3, 10, 14), False),      # norm-and-send kernels + v scatter
    (4, (512, 512, 4, 128), (1, 48, 3, 10, 14), True),
])
def test_every_cross_rank_conflict_is_ordered_by_a_barrier(tmp_path, world, dims, shape, fused):
    mp.spawn(_worker, args=(world, _free_port(), dims, shape, fused, str(tmp_path)), nprocs=world, join=True)
    logs = [torch.load(os.path.join(tmp_path, f"log{r}.pt")) for r in range(world)]
    blocks, forwards = 3, 2
    for log in logs:
        assert [e[1:] for e in log if e[0] == "bar"] == [(s, e + 1) for e in range(blocks * forwards) for s in (0, 1)]
        names = {e[1] for e in log if e[0] == "acc"}
        assert {"attention_scatter", "gemm_o", "attn_cross", "gemm_cross_o"} <= names, names      # the recorder saw the consumers of `o`
        assert ("gemm_qkv_scatter" in names) == fused and ("rmsnorm_rope_scatter" in names) == (not fused)
    assert hazards(logs) == []
    # the checker is not vacuous: without barrier 1 of some block the next send races with the attention still reading `recv`
    # (and the next attention's stores with the consumers of `o`); without barrier 0 the attention reads rows not yet sent
    for which, nth in ((1, 2), (0, 3), (1, blocks * forwards - 1)):
        found = hazards(_without_barrier(logs, which, nth))
        assert found, (which, nth)
        regions = {f[2].split(": ")[1].split(" ")[0] for f in found}
        assert "recv" in regions, (which, nth, regions)


def _train_worker(rank, world, port, shape, recompute, out_dir):
    """The stage-2 training step of tests/test_engine_p2p_host.py with the recorder on.  The trainer calls the kernels directly, so
    every entry point of `ops` is wrapped to log the arena memory its tensor arguments cover, and so is Tensor.copy_ (the trainer
    saves `recv` before barrier 1 and `o` after it with plain copies)."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import types

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
    spmod.PeerArena._allocate = staticmethod(shm.allocate)
    par = fg.SequenceParallel(exchange="p2p")
    log = []
    _recording_exchange(ops, shm, log, par)

    def logged(name, fn):
        def call(*args, **kwargs):
            ar = par.arena
            if ar is not None:
                rects = [r for a in list(args) + list(kwargs.values()) if torch.is_tensor(a) for r in _rects(ar, a)]
                if rects:
                    log.append(("acc", name, rects))
            return fn(*args, **kwargs)
        return call

    for name in dir(ops):
        fn = getattr(ops, name)
        if isinstance(fn, types.FunctionType) and not name.startswith("_"):
            setattr(ops, name, logged(name, fn))
    plain_copy = torch.Tensor.copy_

    def copy_(self, src, *args, **kwargs):
        ar = par.arena
        if ar is not None and torch.is_tensor(src):
            rects = _rects(ar, src) + _rects(ar, self)
            if rects:
                log.append(("acc", "Tensor.copy_", rects))
        return plain_copy(self, src, *args, **kwargs)

    torch.Tensor.copy_ = copy_
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w, lora = o.make_weights(o.TINY, seed=0), o.make_lora(o.TINY, rank=32, seed=2)
    eng = _bare_engine(fg, ops, cfg, par)
    eng.load_state_dict(w)
    tr = Stage2Trainer(eng, lora, rank=32, stage=2, recompute=recompute)
    holder["trainer"] = tr
    tr.load_b2(t.make_b2(o.TINY, rank=32))
    x0, _, ctx, _ = o.make_inputs(o.TINY, shape, text_len=32, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    for step in range(2):                                                    # two steps: the second forward follows the first backward
        tr.zero_grad()
        tr.step(x0, noise, 500, ctx, masks=t.make_masks(o.TINY, rank=32))
    par.check()
    torch.Tensor.copy_ = plain_copy
    torch.save(log, os.path.join(out_dir, f"log{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("recompute", [False, True])
def test_training_exchange_forward_and_backward_is_race_free(tmp_path, recompute):
    world, shape = 2, (1, 48, 3, 10, 14)
    mp.spawn(_train_worker, args=(world, _free_port(), shape, recompute, str(tmp_path)), nprocs=world, join=True)
    logs = [torch.load(os.path.join(tmp_path, f"log{r}.pt")) for r in range(world)]
    exchanges = 2 * 2 * (3 if recompute else 2)                               # steps x blocks x (forward [+ re-computed forward] + backward)
    for log in logs:
        assert [e[1:] for e in log if e[0] == "bar"] == [(s, e + 1) for e in range(exchanges) for s in (0, 1)]
        names = {e[1] for e in log if e[0] == "acc"}
        assert {"attention_scatter", "sp_return_heads", "attention_bwd", "Tensor.copy_", "sp_scatter_heads"} <= names, names
        assert any(r[1] == "dqkv" and r[0] == logs.index(log) for e in log if e[0] == "acc" for r in e[2])   # its consumers were seen
    assert hazards(logs) == []
    for which, nth in ((1, 1), (0, 4), (1, exchanges - 1)):
        assert hazards(_without_barrier(logs, which, nth)), (which, nth)
