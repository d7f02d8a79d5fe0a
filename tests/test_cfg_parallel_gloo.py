"""CPU (gloo) tests of the shot x CFG x Ulysses layout in fairygen_b200.cfg_parallel: rank arithmetic, group
construction and the guidance-pair exchange run for real on 4 processes; the DiT forward is a stand-in
(the kernels need a GPU), so what is checked is who computes what and that every rank ends with
(noise_pos, noise_neg) of ITS shot — the sequential reference order of wan_video.py:296-301."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from fairygen_b200.cfg_parallel import Layout  # noqa: E402


def test_layout_coordinates():
    lay = Layout(8, shots=2, cfg=2, sp=2)
    assert [lay.coords(r) for r in range(8)] == [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1),
                                                 (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]
    assert lay.sp_ranks(5) == [4, 5] and lay.cfg_ranks(5) == [5, 7]
    assert lay.shots_of(6, 5) == [1, 3] and lay.shots_of(0, 5) == [0, 2, 4]
    with pytest.raises(ValueError):
        Layout(8, shots=3, cfg=2, sp=1)
    with pytest.raises(ValueError):
        Layout(8, shots=1, cfg=4, sp=2)


def test_layout_auto_prefers_communication_free_axes():
    assert Layout.auto(8, n_shots=4) == Layout(8, 4, 2, 1)          # BASELINE config 4
    assert Layout.auto(8, n_shots=1) == Layout(8, 1, 2, 4)
    assert Layout.auto(8, n_shots=1, cfg_on=False) == Layout(8, 1, 1, 8)
    assert Layout.auto(4, n_shots=3) == Layout(4, 2, 2, 1)
    assert Layout.auto(1, n_shots=4) == Layout(1, 1, 1, 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeEngine:
    """forward = latents * 2 + mean(context) + timestep/1000 (+ the SP group must see all its ranks)."""

    def __init__(self, sp):
        self.sp = sp
        self.calls = []

    def forward(self, latents, timestep, context, fuse=True):
        self.calls.append(float(context.mean()))
        out = latents * 2 + context.mean() + timestep.reshape(-1)[0] / 1000
        if self.sp is not None:  # every rank of the Ulysses group takes part in each forward
            t = torch.ones(1)
            dist.all_reduce(t, group=self.sp.group)
            assert int(t.item()) == self.sp.world
        return out


def _worker(rank, world, port, shots, cfg, sp, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fairygen_b200.cfg_parallel import Layout, ParallelContext, denoise_shots

    lay = Layout(world, shots, cfg, sp)
    ctx = ParallelContext(lay, exchange="nccl")
    eng = _FakeEngine(ctx.sequence_parallel())
    n_shots = 3
    data = []
    for i in range(n_shots):
        g = torch.Generator().manual_seed(100 + i)
        data.append(dict(latents=torch.randn(1, 4, 2, 4, 4, generator=g), context_pos=torch.randn(1, 8, 16, generator=g) + 1,
                         context_neg=torch.randn(1, 8, 16, generator=g) - 1, first_frame_latents=None))

    class FakeDenoiser:
        cfg_group = None

        def __call__(self, latents, cp, cn, z0):
            lat = latents.clone()
            for step in range(2):
                ts = torch.tensor([1000.0 - 4 * step])
                if self.cfg_group is not None:
                    npos, nneg = self.cfg_group.forward_pair(eng, lat, ts, cp, cn, False)
                else:
                    npos, nneg = eng.forward(lat, ts, cp), eng.forward(lat, ts, cn)
                lat = lat + (nneg + 5.0 * (npos - nneg)) * -0.1
            return lat

    done = denoise_shots(ctx, FakeDenoiser, data)
    # sequential single-process answer
    plain = _FakeEngine(None)
    errs = []
    for i, lat in done:
        ref = data[i]["latents"].clone()
        for step in range(2):
            ts = torch.tensor([1000.0 - 4 * step])
            npos, nneg = plain.forward(ref, ts, data[i]["context_pos"]), plain.forward(ref, ts, data[i]["context_neg"])
            ref = ref + (nneg + 5.0 * (npos - nneg)) * -0.1
        errs.append(float((lat - ref).abs().max()))
    torch.save({"shots": [i for i, _ in done], "errs": errs, "forwards": len(eng.calls), "coords": lay.coords(rank)},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("shots,cfg,sp", [(1, 2, 2), (2, 2, 1), (2, 1, 2)])
def test_layout_world4(tmp_path, shots, cfg, sp):
    world = 4
    mp.spawn(_worker, args=(world, _free_port(), shots, cfg, sp, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        shot_group = res["coords"][0]
        assert res["shots"] == list(range(shot_group, 3, shots)), res
        assert all(e < 1e-5 for e in res["errs"]), res
        # CFG-parallel halves the forwards each rank runs (2 steps per shot)
        assert res["forwards"] == len(res["shots"]) * 2 * (1 if cfg == 2 else 2), res
