"""Ulysses sequence parallelism on real GPUs (NCCL over NVLink): 2 ranks must reproduce the single-GPU
forward (our SP masks padded keys, so equality holds for ragged S too).  Skipped with fewer than 2 GPUs."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BF = torch.bfloat16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, exchange, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", rank))
    cfg = fg.WanDiTConfig(dim=1024, ffn_dim=2048, text_dim=256, num_heads=8, num_layers=2)
    sd = synthetic.random_state_dict(cfg, seed=0, device=f"cuda:{rank}", dtype=BF)
    lat, z0, cp, cn = synthetic.synthetic_inputs(cfg, shape, text_len=64, live_text=16, pin=False)
    ts = torch.tensor([996.0])
    single = fg.WanDiTEngine(cfg, f"cuda:{rank}")
    single.load_state_dict(sd)
    ref = single.forward(lat.cuda(), ts, cp.cuda(), True)
    par = fg.WanDiTEngine(cfg, f"cuda:{rank}", sp=fg.SequenceParallel(exchange=exchange))
    par.load_state_dict(sd)
    out = par.forward(lat.cuda(), ts, cp.cuda(), True)
    fg.ops.sync_check()
    err = float((out.float() - ref.float()).norm() / ref.float().norm())
    torch.save({"err": err, "finite": bool(torch.isfinite(out.float()).all())}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])   # NVLink peer stores from our kernels / NCCL all-to-all
@pytest.mark.parametrize("shape", [(1, 48, 4, 16, 16), (1, 48, 3, 10, 14)])   # S = 256 (even split) and S = 105 (ragged)
def test_sp2_equals_single_gpu(tmp_path, shape, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), shape, exchange, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["finite"] and res["err"] < 3e-3, res   # same kernels; only the attention work split / key-split tail differs


def _oracle_worker(rank, world, port, dims, shape, out_dir):
    """Ulysses SP over `world` GPUs against the ORACLE (the reference's op chain, bf16, run by torch on this rank's GPU) — and
    against the unmodified reference where baseline/_ref travelled to the box — not against our own single-GPU engine."""
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic
    from oracle import wan_dit_oracle as o

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    dim, ffn, heads, text_dim = dims
    cfg = fg.WanDiTConfig(dim=dim, ffn_dim=ffn, text_dim=text_dim, num_heads=heads, num_layers=2)
    ocfg = o.DiTConfig(dim=dim, ffn_dim=ffn, text_dim=text_dim, num_heads=heads, num_layers=2)
    sd = synthetic.random_state_dict(cfg, seed=0, device=dev, dtype=BF, lora_rank=32)
    lat, z0, cp, cn = synthetic.synthetic_inputs(cfg, shape, text_len=128, live_text=24, pin=False)
    lat, cp = lat.to(dev), cp.to(dev)
    ts = torch.tensor([996.0], device=dev, dtype=BF)
    par = fg.WanDiTEngine(cfg, dev, sp=fg.SequenceParallel(exchange="p2p"))
    par.load_state_dict(sd)
    out = par.forward(lat, ts, cp, True)
    fg.ops.sync_check()
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())  # noqa: E731
    with torch.no_grad():
        ref16 = o.dit_forward(sd, ocfg, lat, ts, cp, True)
        ref32 = o.dit_forward({k: v.float() for k, v in sd.items()}, ocfg, lat.float(), ts.float(), cp.float(), True)
    res = {"err_oracle_bf16": rel(out, ref16), "err_oracle_fp32": rel(out, ref32), "ref_bf16_vs_fp32": rel(ref16, ref32),
           "finite": bool(torch.isfinite(out.float()).all()), "err_reference": None}
    from baseline import ref_loader as rl

    if rl.available():
        ref = rl.load()
        rl.select_attention_backend(dev)
        dit = rl.build_wan_model(cfg, state_dict=sd)
        with torch.no_grad():
            want = ref.wv.model_fn_wan_video(dit=dit, latents=lat, timestep=ts, context=cp, fuse_vae_embedding_in_latents=True)
        res["err_reference"] = rel(out, want)
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("dims,shape", [
    ((3072, 14336, 24, 4096), (1, 48, 3, 10, 14)),    # TI2V-5B block shapes, 12 heads per rank, S = 105 (ragged: 1 pad row)
    ((3072, 14336, 24, 4096), (1, 48, 5, 16, 16)),    # S = 320 (even split)
    ((768, 2048, 6, 512), (1, 48, 3, 10, 14)),        # 3 heads per rank — the per-rank head count of Ulysses SP8 on TI2V-5B
    # S = 1400, 700 rows per rank: the q|k|v GEMM-with-send has 3 x 36 = 108 tiles = one wave of 74 CTA pairs + 34, so its last
    # wave runs as 68 half-width items whose 64-column blocks still go to the right head owners
    ((3072, 14336, 24, 4096), (1, 48, 5, 28, 40)),
])
def test_sp2_vs_oracle_at_north_star_tolerance(tmp_path, dims, shape):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_oracle_worker, args=(world, _free_port(), dims, shape, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        print(res)
        assert res["finite"] and res["err_oracle_bf16"] < 1e-2, res          # north star: rel L2 <= 1e-2 per forward
        assert res["err_oracle_fp32"] < 2.0 * res["ref_bf16_vs_fp32"] + 1e-3, res   # no worse than the reference's own bf16 error
        assert res["err_reference"] is None or res["err_reference"] < 1e-2, res


def _cfg_worker(rank, world, port, shape, sp_ways, out_dir):
    """CFG-parallel pair (x Ulysses inside each half): a 3-step denoise must equal the single-GPU sequential loop."""
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic
    from fairygen_b200.cfg_parallel import Layout, ParallelContext

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    cfg = fg.WanDiTConfig(dim=1024, ffn_dim=2048, text_dim=256, num_heads=8, num_layers=2)
    sd = synthetic.random_state_dict(cfg, seed=0, device=dev, dtype=BF)
    lat, z0, cp, cn = synthetic.synthetic_inputs(cfg, shape, text_len=64, live_text=16, pin=False)
    single = fg.WanDiTEngine(cfg, dev)
    single.load_state_dict(sd)
    ref = fg.WanDenoiser(single, 4)(lat, cp, cn, z0, steps=range(3))
    par = ParallelContext(Layout(world, 1, 2, sp_ways))
    eng = fg.WanDiTEngine(cfg, dev, sp=par.sequence_parallel())
    eng.load_state_dict(sd)
    out = fg.WanDenoiser(eng, 4, cfg_group=par)(lat, cp, cn, z0, steps=range(3))
    fg.ops.sync_check()
    err = float((out.float() - ref.float()).norm() / ref.float().norm())
    torch.save({"err": err, "finite": bool(torch.isfinite(out.float()).all())}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,sp_ways", [(2, 1), (4, 2)])
def test_cfg_parallel_equals_sequential(tmp_path, world, sp_ways):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_cfg_worker, args=(world, _free_port(), (1, 48, 3, 10, 14), sp_ways, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        # sp_ways == 1: bit-identical kernels on both sides, only the exchange differs
        assert res["finite"] and res["err"] < (1e-6 if sp_ways == 1 else 2e-3), res


def _train_worker(rank, world, port, shape, stage, recompute, out_dir):
    """Sequence-parallel LoRA training step (tokens of one video over 2 GPUs, forward AND backward exchange through peer
    memory) against the single-GPU trainer on the same weights, adapters, dropout masks, noise and timestep."""
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic
    from fairygen_b200.training import Stage2Trainer

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    cfg = fg.WanDiTConfig(dim=512, ffn_dim=1024, text_dim=256, num_heads=4, num_layers=2)
    sd = synthetic.random_state_dict(cfg, seed=0, device=dev, dtype=BF)
    lora = synthetic.random_lora(cfg, rank=32, seed=2, device=dev)
    x0, _, ctx, _ = synthetic.synthetic_inputs(cfg, shape, text_len=32, live_text=8, pin=False)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    gen = torch.Generator().manual_seed(11)
    res = {}
    trainers = []
    for sp in (None, fg.SequenceParallel(exchange="p2p")):
        eng = fg.WanDiTEngine(cfg, dev, sp=sp)
        eng.load_state_dict(sd)
        tr = Stage2Trainer(eng, lora, rank=32, stage=stage, recompute=recompute and sp is not None)
        trainers.append(tr)
    single, par = trainers
    b2 = {t: torch.randn(single.b2[t].shape, generator=gen) * 0.05 for t in single.targets}
    masks = {t: (torch.rand(single.b2[t].shape, generator=gen) > single.dropout_prob).to(torch.uint8) for t in single.targets}
    out = []
    for tr in trainers:
        tr.load_b2(b2)
        tr.zero_grad()
        loss, pred = tr.step(x0, noise, 400, ctx.to(dev), masks=masks, return_pred=True)
        fg.ops.sync_check()
        out.append((float(loss), pred.float().clone(), tr.grad_flat.clone(), None if tr.grad_a_flat is None else tr.grad_a_flat.clone()))
    rel = lambda a, b: float((a - b).norm() / b.norm())  # noqa: E731
    res["loss"] = (out[1][0], out[0][0])
    res["pred"] = rel(out[1][1], out[0][1])
    res["grad"] = rel(out[1][2], out[0][2])
    res["grad_a"] = 0.0 if out[0][3] is None else rel(out[1][3], out[0][3])
    res["worst"] = max(rel(par.grad[t], single.grad[t]) for t in single.targets)
    res["nonzero"] = bool(out[0][2].abs().max() > 0)
    # a second micro-step accumulates on top of the reduced gradients (the reduction covers one backward only)
    par.step(x0, noise, 400, ctx.to(dev), masks=masks)
    fg.ops.sync_check()
    res["accum"] = rel(par.grad_flat, 2 * out[1][2])
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,stage,recompute", [((1, 48, 3, 10, 14), 2, False),    # S = 105: ragged split, one padded row
                                                   ((1, 48, 4, 8, 8), 1, True)])      # S = 64, stage 1 (dA and dB), checkpointing
def test_sp2_training_step_equals_single_gpu(tmp_path, shape, stage, recompute):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_train_worker, args=(world, _free_port(), shape, stage, recompute, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        print(res)
        assert res["nonzero"]
        assert res["pred"] < 3e-3, res
        assert abs(res["loss"][0] - res["loss"][1]) < 5e-3 * abs(res["loss"][1]), res
        assert res["grad"] < 1e-2 and res["grad_a"] < 1e-2 and res["worst"] < 3e-2, res
        assert res["accum"] < 1e-3, res


def _vae_worker(rank, world, port, out_dir):
    """Tiled VAE decode with the windows spread over 2 GPUs == the same decode on one GPU."""
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    from fairygen_b200 import ops, vae
    from oracle import vae38_oracle as o

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    dec = vae.VAE38Decoder(vae.VAE38Config(z_dim=o.TINY.z_dim, dec_dim=o.TINY.dec_dim), dev)
    dec.load_state_dict(o.make_weights(o.TINY, seed=0))
    z = torch.randn((1, o.TINY.z_dim, 2, 6, 7), generator=torch.Generator().manual_seed(3)).to(BF)
    kw = dict(tiled=True, tile_size=(3, 4), tile_stride=(2, 3))
    alone = dec.decode(z, **kw)
    shared = dec.decode(z, group=dist.group.WORLD, **kw)
    ops.sync_check()
    err = float((shared.float() - alone.float()).abs().max())
    torch.save({"err": err, "finite": bool(torch.isfinite(shared.float()).all())}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_vae_windows_over_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_vae_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["finite"] and res["err"] < 1e-2, res       # bf16 output; fp32 sums of 2-4 contributions in another order


def _bcast_worker(rank, world, port, ckpt_pattern, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic, weights

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    eng = fg.WanDiTEngine(cfg, dev)
    if rank != 0:
        def forbidden(*a, **k):
            raise AssertionError("a non-source rank read the checkpoint from disk")
        real_load, weights.load_state_dict = weights.load_state_dict, forbidden
    weights.load_engine_broadcast(eng, ckpt_pattern, src=0)
    if rank != 0:
        weights.load_state_dict = real_load
    direct = fg.WanDiTEngine(cfg, dev)
    direct.load_state_dict(weights.load_state_dict(ckpt_pattern, device=dev))
    same = all(torch.equal(a, b) for a, b in zip(weights.packed_tensors(eng), weights.packed_tensors(direct)))
    lat, z0, cp, cn = synthetic.synthetic_inputs(cfg, (1, 48, 3, 8, 8), text_len=32, live_text=8, pin=False)
    ts = torch.tensor([900.0])
    out = eng.forward(lat.to(dev), ts, cp.to(dev), True)
    ref = direct.forward(lat.to(dev), ts, cp.to(dev), True)
    fg.ops.sync_check()
    torch.save({"same": same, "forward_equal": bool(torch.equal(out, ref))}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_load_engine_broadcast_two_gpus(tmp_path):
    """weights.load_engine_broadcast over NCCL / NVLink: rank 0 reads the sharded checkpoint, rank 1 receives the packed tensors
    (models/model_loader.py:62-80 has every rank read everything)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import safetensors.torch as st
    import torch.multiprocessing as mp

    import fairygen_b200 as fg
    from fairygen_b200.synthetic import param_shapes

    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    g = torch.Generator().manual_seed(3)
    sd = {k: torch.randn(s, generator=g).to(BF) for k, s in param_shapes(cfg).items()}
    names = sorted(sd)
    st.save_file({k: sd[k] for k in names[::2]}, str(tmp_path / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    st.save_file({k: sd[k] for k in names[1::2]}, str(tmp_path / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    mp.spawn(_bcast_worker, args=(2, _free_port(), os.path.join(str(tmp_path), "diffusion_pytorch_model*.safetensors"), str(tmp_path)),
             nprocs=2, join=True)
    for r in range(2):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["same"] and res["forward_equal"], res
