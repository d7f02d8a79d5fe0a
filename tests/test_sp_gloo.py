"""world_size-2 CPU test (gloo) of the Ulysses exchange in fairygen_b200.sp: the product's collective
plumbing runs for real; the three CUDA kernels it calls (pack, attention, unpack) are replaced IN THIS
TEST by torch emulations of their documented layouts, so what is checked is the partition/exchange logic:
rank-local [rows, 3*H*128] -> all-to-all -> full-sequence attention on H/P heads -> all-to-all -> [rows, H*128]
must equal single-rank attention, including a ragged S (padded keys masked)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _emulated_pack(x, send, heads, groups, world):
    rows = x.shape[0]
    hpr = heads // world
    v = x.view(rows, groups, world, hpr, 128).permute(2, 0, 1, 3, 4)  # [world][rows][groups][hpr][128]
    send.view(world, rows, groups, hpr, 128).copy_(v)
    return send


def _emulated_unpack(recv, x, heads, groups, world):
    rows = x.shape[0]
    hpr = heads // world
    v = recv.view(world, rows, groups, hpr, 128).permute(1, 2, 0, 3, 4)
    x.view(rows, groups, world, hpr, 128).copy_(v)
    return x


def _emulated_attention(q, k, v, out, heads, scale=None, lse=None, kmax2=None):
    from oracle import wan_dit_oracle as o
    out.copy_(o.attention(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), heads)[0])
    return out


def _worker(rank, world, port, tokens, heads, result_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fairygen_b200 import sp as spmod
    from oracle import wan_dit_oracle as o

    spmod.ops.sp_pack_heads = _emulated_pack
    spmod.ops.sp_unpack_heads = _emulated_unpack
    spmod.ops.attention = _emulated_attention

    g = torch.Generator().manual_seed(7)
    d = heads * 128
    qkv_full = torch.randn(tokens, 3 * d, generator=g)
    rows, tok0, real = spmod.partition(tokens, world, rank)
    qkv = torch.zeros(rows, 3 * d)
    qkv[:real] = qkv_full[tok0:tok0 + real]
    qkv[real:] = 5.0  # garbage in the padded rows must not leak into real tokens
    s_pad = rows * world
    ws = dict(send=torch.empty(s_pad, 3 * d // world), recv=torch.empty(s_pad, 3 * d // world),
              o_full=torch.empty(s_pad, d // world), o_recv=torch.empty(s_pad, d // world))

    class Eng:
        class cfg:
            num_heads = heads

        @staticmethod
        def _k(name, fn, *a, **kw):
            return fn(*a, **kw)

    par = spmod.SequenceParallel(exchange="nccl")   # the collective variant (gloo here); the NVLink peer-store variant needs GPUs
    assert (par.world, par.rank) == (world, rank)
    out = torch.empty(rows, d)
    par.attention(Eng, ws, qkv, out, tokens)
    full = o.attention(qkv_full[None, :, :d], qkv_full[None, :, d:2 * d], qkv_full[None, :, 2 * d:], heads)[0]
    err = (out[:real] - full[tok0:tok0 + real]).abs().max().item()
    gathered = torch.empty(s_pad, d)
    par.all_gather_rows(out, gathered)
    err2 = (gathered[:tokens] - full).abs().max().item() if rank == 0 else 0.0
    torch.save({"err": err, "err2": err2}, os.path.join(result_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("tokens", [64, 45])
def test_ulysses_exchange_world2(tmp_path, tokens):
    world, heads = 2, 4
    mp.spawn(_worker, args=(world, _free_port(), tokens, heads, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["err"] < 1e-5 and res["err2"] < 1e-5, res
