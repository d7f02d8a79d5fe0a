"""world_size-2 CPU test (gloo) of the WHOLE DiT forward under Ulysses sequence parallelism: the product's engine, token
partition, per-rank first-frame rows, RoPE token offsets, exchange (NCCL-variant plumbing: pack / all-to-all / unpack, here over
gloo) and output gather run for real; the kernels are replaced by the plain-torch contract statements of
tests/test_engine_host.py.  Each rank's gathered prediction must equal the single-process forward and the oracle, for an even
split and for a ragged one (S = 105 tokens over 2 ranks: one padded row, masked as a key)."""
import os
import socket
import sys

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
    """The two methods of pytest's monkeypatch that _emulated_ops uses (the worker runs outside pytest)."""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def _bare_engine(fg, ops, cfg, sp):
    eng = fg.WanDiTEngine.__new__(fg.WanDiTEngine)          # the constructor refuses the CPU (tests/test_engine_host.py)
    eng.cfg, eng.device, eng.ctx, eng.sp = cfg, torch.device("cpu"), None, sp
    eng.rope_tab = torch.from_numpy(ops.rope_table(cfg.head_dim))
    eng._init_state()
    return eng


def _worker(rank, world, port, shape, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)      # several processes of many tiny CPU ops: the default thread count oversubscribes the cores badly
    import fairygen_b200 as fg
    from fairygen_b200 import ops
    from oracle import wan_dit_oracle as o
    from test_engine_host import _emulated_ops
    from test_sp_gloo import _emulated_pack, _emulated_unpack

    _emulated_ops(_Patch)
    ops.sp_pack_heads, ops.sp_unpack_heads = _emulated_pack, _emulated_unpack
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w = o.make_weights(o.TINY, seed=0)
    lat, _, cp, _ = o.make_inputs(o.TINY, shape, text_len=32, live_text=8)
    ts = torch.tensor([900.0])
    single = _bare_engine(fg, ops, cfg, None)
    single.load_state_dict(w)
    ref = single.forward(lat.to(BF), ts, cp.to(BF), True)
    par = _bare_engine(fg, ops, cfg, fg.SequenceParallel(exchange="nccl"))
    par.load_state_dict(w)
    out = par.forward(lat.to(BF), ts, cp.to(BF), True)
    with torch.no_grad():
        want = o.dit_forward({k: v.to(BF).float() for k, v in w.items()}, o.TINY, lat.to(BF).float(), ts, cp.to(BF).float(), True)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())  # noqa: E731
    torch.save({"vs_single": rel(out.float(), ref.float()), "vs_oracle": rel(out.float(), want), "shape": tuple(out.shape)},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(1, 48, 4, 8, 8), (1, 48, 3, 10, 14)])       # S = 64 (32 + 32) and S = 105 (53 + 52 and a padded row)
def test_sequence_parallel_forward_world2(tmp_path, shape):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), shape, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["shape"] == shape
        assert res["vs_single"] < 4e-3, res          # same contracts; only the summation order inside attention differs
        assert res["vs_oracle"] < 1e-2, res


def _layout_worker(rank, world, port, shots, cfg_ways, sp_ways, out_dir):
    """Shot x CFG-pair x Ulysses layout with the REAL engine and denoiser on 4 gloo processes: a 2-step CFG denoise of every
    shot must equal the sequential single-process loop (PIPE:285-309; batch_inference.py:45-56)."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)      # several processes of many tiny CPU ops: the default thread count oversubscribes the cores badly
    import fairygen_b200 as fg
    from fairygen_b200 import ops, scheduler
    from fairygen_b200.cfg_parallel import Layout, ParallelContext, denoise_shots
    from oracle import wan_dit_oracle as o
    from test_engine_host import _emulated_ops
    from test_sp_gloo import _emulated_pack, _emulated_unpack

    _emulated_ops(_Patch)
    ops.sp_pack_heads, ops.sp_unpack_heads = _emulated_pack, _emulated_unpack

    def step_fused(self, latents, noise_pos, noise_neg, cfg_scale, index, first_frame_latents=None, to_final=False):
        ops.cfg_fm_step(latents, noise_pos, noise_neg, first_frame_latents, float(cfg_scale), self.sigma_delta(index, to_final))
        return latents

    scheduler.FlowMatchScheduler.step_fused = step_fused          # without the CUDA-only guard (tests/test_host_logic.py checks it)
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    w = o.make_weights(o.TINY, seed=0)
    data = []
    for i in range(2):
        lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 2, 6, 10), text_len=24, live_text=8)
        g = torch.Generator().manual_seed(50 + i)
        data.append(dict(latents=lat + 0.1 * torch.randn(lat.shape, generator=g), context_pos=cp, context_neg=cn, first_frame_latents=z0))
    ctx = ParallelContext(Layout(world, shots, cfg_ways, sp_ways), exchange="nccl")
    eng = _bare_engine(fg, ops, cfg, ctx.sequence_parallel())
    eng.load_state_dict(w)
    done = denoise_shots(ctx, lambda: fg.WanDenoiser(eng, 2, cfg_scale=5.0, sigma_shift=5.0), data)
    single = _bare_engine(fg, ops, cfg, None)
    single.load_state_dict(w)
    errs = []
    for i, lat in done:
        s = data[i]
        ref = fg.WanDenoiser(single, 2, cfg_scale=5.0, sigma_shift=5.0)(s["latents"], s["context_pos"], s["context_neg"], s["first_frame_latents"])
        errs.append(float((lat.double() - ref.double()).norm() / ref.double().norm()))
    torch.save({"shots": [i for i, _ in done], "errs": errs}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("shots,cfg_ways,sp_ways", [(1, 2, 2), (2, 2, 1)])
def test_shot_cfg_sp_layout_world4(tmp_path, shots, cfg_ways, sp_ways):
    world = 4
    mp.spawn(_layout_worker, args=(world, _free_port(), shots, cfg_ways, sp_ways, str(tmp_path)), nprocs=world, join=True)
    seen = set()
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        seen.update(res["shots"])
        assert res["shots"] and all(e < 5e-3 for e in res["errs"]), res
    assert seen == {0, 1}


def _vae_worker(rank, world, port, out_dir):
    """VAE38 tiled decode with the windows spread over 2 gloo ranks (assign_windows + one sum all-reduce before the blend)."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)      # several processes of many tiny CPU ops: the default thread count oversubscribes the cores badly
    from fairygen_b200 import vae
    from oracle import vae38_oracle as o
    from test_vae_host import _emulated_ops

    _emulated_ops(_Patch)
    dec = vae.VAE38Decoder(vae.VAE38Config(z_dim=o.TINY.z_dim, dec_dim=o.TINY.dec_dim), "cpu")
    dec.load_state_dict(o.make_weights(o.TINY, seed=0))
    z = torch.randn((1, o.TINY.z_dim, 2, 6, 7), generator=torch.Generator().manual_seed(3)).to(BF)
    kw = dict(tiled=True, tile_size=(3, 4), tile_stride=(2, 3))
    alone = dec.decode(z, **kw)
    shared = dec.decode(z, group=dist.group.WORLD, **kw)
    torch.save({"err": float((shared.float() - alone.float()).abs().max())}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_vae_windows_over_two_ranks(tmp_path):
    mp.spawn(_vae_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert torch.load(os.path.join(tmp_path, f"r{r}.pt"))["err"] < 1e-2
