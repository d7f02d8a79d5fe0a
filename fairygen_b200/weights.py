"""Checkpoint detection and loading for the DiT.

The reference finds out what a checkpoint is by hashing its sorted ``key:shape`` list (``hash_model_file``,
animation/diffsynth/core/loader/file.py:40-121) and looking the md5 up in ``MODEL_CONFIGS``
(configs/model_configs.py; Wan2.2-TI2V-5B = ``1f5ab7703c6fc803fdded85ff040c316``, :290-295), then every rank of a
multi-GPU job reads the full 10 GB from disk (models/model_loader.py:62-80, core/vram/disk_map.py:28-93).

Here: the same hash (so the same files are recognised — ``param_shapes(TI2V_5B)`` reproduces the reference's hash, which
pins the complete key / shape inventory of the model), shards are read through ``safetensors.safe_open`` (mmap, one
tensor at a time, straight to the GPU), and in a multi-GPU job ONE rank reads and packs while the others receive the
packed bf16 tensors by broadcast over NVLink (``load_engine_broadcast``) — one disk read per box instead of eight.
"""
from __future__ import annotations

import glob
import hashlib
from typing import Dict, Iterable, List, Sequence, Union

import torch

from .config import TI2V_5B, WanDiTConfig

TI2V_5B_HASH = "1f5ab7703c6fc803fdded85ff040c316"   # configs/model_configs.py:291


def keys_hash(shapes: Dict[str, Sequence[int]], with_shape: bool = True) -> str:
    """md5 of the sorted ``key:shape,key`` list — convert_keys_dict_to_single_str + hash_model_file (file.py:99-121)."""
    keys: List[str] = []
    for k, shp in shapes.items():
        if with_shape:
            keys.append(k + ":" + "_".join(map(str, list(shp))))
        keys.append(k)
    keys.sort()
    return hashlib.md5(",".join(keys).encode("utf-8")).hexdigest()


def expand(paths: Union[str, Iterable[str]]) -> List[str]:
    """A path, a glob (``diffusion_pytorch_model*.safetensors``) or a list of them -> sorted shard files."""
    items = [paths] if isinstance(paths, str) else list(paths)
    out: List[str] = []
    for p in items:
        hit = sorted(glob.glob(p)) if any(c in p for c in "*?[") else [p]
        out.extend(hit)
    if not out:
        raise FileNotFoundError(f"no checkpoint files match {paths!r}")
    return out


def file_shapes(paths: Union[str, Iterable[str]]) -> Dict[str, List[int]]:
    """key -> shape over all shards without reading tensor data (load_keys_dict_from_safetensors, file.py:79-84)."""
    from safetensors import safe_open

    shapes: Dict[str, List[int]] = {}
    for path in expand(paths):
        with safe_open(path, framework="pt", device="cpu") as f:
            for k in f.keys():
                shapes[k] = list(f.get_slice(k).get_shape())
    return shapes


def detect(paths: Union[str, Iterable[str]]) -> WanDiTConfig:
    """The config of a checkpoint, by the reference's hash; raises like ModelPool.auto_load_model (model_loader.py:80)."""
    h = keys_hash(file_shapes(paths))
    if h == TI2V_5B_HASH:
        return TI2V_5B
    raise ValueError(f"Cannot detect the model type. File: {paths}. Model hash: {h} (this build knows Wan2.2-TI2V-5B = {TI2V_5B_HASH})")


def load_state_dict(paths: Union[str, Iterable[str]], device="cpu", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """All tensors of the shards on `device` in `dtype` (floating tensors only are cast)."""
    from safetensors import safe_open

    sd: Dict[str, torch.Tensor] = {}
    dev = str(torch.device(device))
    for path in expand(paths):
        with safe_open(path, framework="pt", device=dev) as f:
            for k in f.keys():
                t = f.get_tensor(k)
                sd[k] = t.to(dtype) if t.is_floating_point() else t
    return sd


def packed_tensors(engine) -> List[torch.Tensor]:
    """Every packed weight tensor of a WanDiTEngine in a fixed order (the broadcast schedule)."""
    names = ["w_patch", "b_patch", "w_text0", "b_text0", "w_text2", "b_text2", "w_time0", "b_time0", "w_time2", "b_time2", "w_tproj",
             "b_tproj", "w_head", "b_head", "head_mod", "mods_all"]
    out = [getattr(engine, n) for n in names]
    for b in engine.blocks:
        out.extend(getattr(b, n) for n in b.__slots__ if n != "private")   # `private` is the block's copy-on-write bookkeeping
    return out


def load_engine_broadcast(engine, paths: Union[str, Iterable[str]], group=None, src: int = 0) -> None:
    """Rank `src` of `group` reads + packs the checkpoint; the other ranks allocate the packed layout (from a state dict
    of empty tensors, no disk access) and receive every packed tensor by NCCL broadcast over NVLink."""
    import torch.distributed as dist

    from .synthetic import param_shapes

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        engine.load_state_dict(load_state_dict(paths, device=engine.device))
        return
    if dist.get_rank(group) == src:
        engine.load_state_dict(load_state_dict(paths, device=engine.device))
    else:
        engine.load_state_dict({k: torch.empty(s, dtype=torch.bfloat16, device=engine.device) for k, s in param_shapes(engine.cfg).items()})
    global_src = dist.get_global_rank(group, src) if group is not None else src
    for t in packed_tensors(engine):
        dist.broadcast(t, src=global_src, group=group)
