"""Checkpoint detection / loading (fairygen_b200.weights) against the reference's conventions.  CPU only."""
import os

import pytest
import torch

from fairygen_b200 import TI2V_5B, WanDiTConfig, weights
from fairygen_b200.synthetic import param_shapes


def test_parameter_inventory_reproduces_the_reference_model_hash():
    """KNOWN ANSWER from the reference: configs/model_configs.py:291 lists md5 1f5ab7703c6fc803fdded85ff040c316 for the
    Wan2.2-TI2V-5B checkpoint (hash of its sorted key:shape list, core/loader/file.py:99-121). Our parameter inventory —
    the names and shapes the engine packs from and the synthetic weights are generated for — must hash to the same value,
    i.e. all 825 tensors of the model are accounted for with the right shapes."""
    shapes = param_shapes(TI2V_5B)
    assert len(shapes) == 825
    assert weights.keys_hash(shapes) == weights.TI2V_5B_HASH == "1f5ab7703c6fc803fdded85ff040c316"
    assert sum(int(torch.tensor(s).prod()) for s in shapes.values()) == 4_999_787_712     # parameter count


def test_detect_and_load_sharded_safetensors(tmp_path):
    import safetensors.torch as st
    cfg = WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    shapes = param_shapes(cfg)
    g = torch.Generator().manual_seed(0)
    sd = {k: torch.randn(s, generator=g).to(torch.bfloat16) for k, s in shapes.items()}
    names = sorted(sd)
    st.save_file({k: sd[k] for k in names[::2]}, str(tmp_path / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    st.save_file({k: sd[k] for k in names[1::2]}, str(tmp_path / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    pattern = os.path.join(str(tmp_path), "diffusion_pytorch_model*.safetensors")
    assert len(weights.expand(pattern)) == 2
    assert weights.file_shapes(pattern) == {k: list(v) for k, v in shapes.items()}
    assert weights.keys_hash(weights.file_shapes(pattern)) == weights.keys_hash(shapes)
    with pytest.raises(ValueError, match="Cannot detect the model type"):      # a tiny model is not a known checkpoint
        weights.detect(pattern)
    back = weights.load_state_dict(pattern)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        weights.expand(os.path.join(str(tmp_path), "nothing*.safetensors"))


def _broadcast_worker(rank, world, port, ckpt_pattern, out_dir):
    """Rank 0 reads + packs the checkpoint, rank 1 never touches the files: it allocates the packed layout and receives every
    packed tensor by broadcast (weights.load_engine_broadcast; the reference makes every rank read the full checkpoint,
    models/model_loader.py:62-80, core/vram/disk_map.py:28-93)."""
    import sys

    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fairygen_b200 as fg

    cfg = WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    eng = fg.WanDiTEngine.__new__(fg.WanDiTEngine)      # the constructor refuses the CPU (tests/test_engine_host.py); host logic only
    eng.cfg, eng.device, eng.sp = cfg, torch.device("cpu"), None
    eng._init_state()
    if rank != 0:
        real_load = weights.load_state_dict

        def forbidden(*a, **k):
            raise AssertionError("a non-source rank read the checkpoint from disk")

        weights.load_state_dict = forbidden
    weights.load_engine_broadcast(eng, ckpt_pattern, group=None, src=0)
    if rank != 0:
        weights.load_state_dict = real_load
    want = fg.WanDiTEngine.__new__(fg.WanDiTEngine)
    want.cfg, want.device, want.sp = cfg, torch.device("cpu"), None
    want._init_state()
    want.load_state_dict(weights.load_state_dict(ckpt_pattern))
    same = all(torch.equal(a, b) for a, b in zip(weights.packed_tensors(eng), weights.packed_tensors(want)))
    torch.save({"same": same, "n": len(weights.packed_tensors(eng)), "loaded": eng.loaded}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_load_engine_broadcast_over_two_ranks(tmp_path):
    """f4 of SURVEY §8(f): load once + broadcast. world-2 gloo (the GPU path is the same code over NCCL,
    tests/test_sp_gpu.py::test_load_engine_broadcast_two_gpus)."""
    import socket

    import safetensors.torch as st
    import torch.multiprocessing as mp

    cfg = WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    g = torch.Generator().manual_seed(3)
    sd = {k: torch.randn(s, generator=g).to(torch.bfloat16) for k, s in param_shapes(cfg).items()}
    names = sorted(sd)
    st.save_file({k: sd[k] for k in names[::2]}, str(tmp_path / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    st.save_file({k: sd[k] for k in names[1::2]}, str(tmp_path / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_broadcast_worker, args=(2, port, os.path.join(str(tmp_path), "diffusion_pytorch_model*.safetensors"), str(tmp_path)),
             nprocs=2, join=True)
    for r in range(2):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["same"] and res["loaded"] and res["n"] == 16 + 2 * 20, res
