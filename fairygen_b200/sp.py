"""Ulysses sequence parallelism for the DiT self-attention (one process per GPU, torch.distributed).

Semantics follow the reference's USP glue (animation/diffsynth/utils/xfuser/xdit_context_parallel.py
:11-21, 125-146 and pipelines/wan_video.py:1224-1227, 1310-1315, 1379-1382): tokens are split
contiguously in (f h w) order into ``world`` chunks of ceil(S/world) rows (zero-padded tail),
weights are replicated, and inside self-attention the layout is swapped to "all tokens, heads/world
heads" by an all-to-all and swapped back afterwards.  The all-to-all itself lives in ``xfuser`` in
the reference (not in its tree, unpinned); here it is NCCL over NVLink on buffers laid out by the
``fgb_sp_pack_heads`` / ``fgb_sp_unpack_heads`` kernels so that the receive buffer is directly the
[tokens, heads*128] matrix the attention kernel reads.  Unlike the reference, padded keys are
masked (``s_kv = S``), so the result equals the single-GPU result for any S.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

from . import ops


def partition(tokens: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(rows per rank, first global token of `rank`, number of real tokens on `rank`)."""
    rows = -(-tokens // world)
    tok0 = rank * rows
    return rows, tok0, max(0, min(rows, tokens - tok0))


def first_frame_rows(tokens_first_frame: int, world: int, rank: int, tokens: int) -> int:
    """How many of this rank's rows belong to the first latent frame (timestep 0, PIPE:1218-1222)."""
    rows, tok0, _ = partition(tokens, world, rank)
    return max(0, min(rows, tokens_first_frame - tok0))


def init_process_group_from_env(device_type: str = "cuda"):
    """Mirror of initialize_usp (xdit_context_parallel.py:11-21) without xfuser: env:// rendezvous,
    NCCL on GPUs (gloo on CPU for the host-logic tests), one device per local rank."""
    import os

    if not dist.is_initialized():
        dist.init_process_group(backend="nccl" if device_type == "cuda" else "gloo", init_method="env://")
    if device_type == "cuda":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", dist.get_rank())))
    return dist.group.WORLD


class SequenceParallel:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("SequenceParallel needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)

    # ---- collectives (thin: NCCL on device tensors, gloo in the CPU tests) ----------------------
    def all_to_all(self, recv: torch.Tensor, send: torch.Tensor) -> torch.Tensor:
        dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def all_gather_rows(self, rows: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        dist.all_gather_into_tensor(out, rows.contiguous(), group=self.group)
        return out

    # ---- the self-attention exchange ---------------------------------------------------------------
    def attention(self, engine, ws, qkv: torch.Tensor, o: torch.Tensor, tokens: int) -> None:
        """qkv [rows, 3*H*128] (q,k already RMS-normed + rotated) -> o [rows, H*128]."""
        heads = engine.cfg.num_heads
        hpr = heads // self.world
        wloc = hpr * 128
        k = engine._k
        k("sp_pack", ops.sp_pack_heads, qkv, ws["send"], heads, 3, self.world)
        recv = k("sp_all_to_all", self.all_to_all, ws["recv"], ws["send"])     # [s_pad tokens, (q|k|v) x hpr x 128]
        k("attn_self", ops.attention, recv[:, :wloc], recv[:tokens, wloc:2 * wloc], recv[:tokens, 2 * wloc:], ws["o_full"], hpr)
        o_recv = k("sp_all_to_all", self.all_to_all, ws["o_recv"], ws["o_full"])  # [world][rows][hpr*128]
        k("sp_unpack", ops.sp_unpack_heads, o_recv, o, heads, 1, self.world)
