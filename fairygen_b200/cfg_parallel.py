"""Process-group layout for one 8-GPU box: shot-parallel x CFG-parallel x Ulysses sequence-parallel.

The reference runs classifier-free guidance as two *sequential* batch-1 DiT forwards per step
(animation/diffsynth/pipelines/wan_video.py:296-304; ``cfg_merge`` at :785-803 only batches them) and
``batch_inference.py:45-52`` renders the shots of a story one after the other in one process.  Both
loops are embarrassingly parallel, so SURVEY.md §8(e) adds two outer axes around the Ulysses group:

    rank = (shot_group * cfg_ways + cfg_rank) * sp_ways + sp_rank

  * ``sp`` (innermost, NVLink all-to-all per attention): ranks that split ONE forward's tokens;
  * ``cfg`` (2 ways): group 0 runs the positive prompt, group 1 the negative prompt of the same step;
    afterwards each rank exchanges its noise prediction with its partner (one 2-rank all-gather of
    the 10.5 MB prediction) and every rank applies CFG + Euler + first-frame restore locally
    (``fgb_cfg_fm_step``), so latents stay replicated and no broadcast is needed;
  * ``shot`` (outermost): independent videos (shots of ``batch_inference``), no communication.

Config 4 of BASELINE.json (4 shots 480x832x81 on 8 GPUs) is ``Layout(world=8, shots=4, cfg=2, sp=1)``
or ``Layout(world=8, shots=2, cfg=2, sp=2)``; the headline single video at 8 GPUs is either
``Layout(8, 1, 1, 8)`` (pure Ulysses) or ``Layout(8, 1, 2, 4)``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Layout:
    world: int
    shots: int = 1
    cfg: int = 1
    sp: int = 1

    def __post_init__(self):
        if self.cfg not in (1, 2):
            raise ValueError(f"cfg ways must be 1 or 2 (positive / negative prompt), got {self.cfg}")
        if self.shots < 1 or self.sp < 1 or self.shots * self.cfg * self.sp != self.world:
            raise ValueError(f"layout shots={self.shots} x cfg={self.cfg} x sp={self.sp} != world {self.world}")

    def coords(self, rank: int) -> Tuple[int, int, int]:
        """(shot_group, cfg_rank, sp_rank) of a global rank."""
        if not 0 <= rank < self.world:
            raise ValueError(f"rank {rank} outside world {self.world}")
        return rank // (self.cfg * self.sp), (rank // self.sp) % self.cfg, rank % self.sp

    def sp_ranks(self, rank: int) -> List[int]:
        base = rank - rank % self.sp
        return list(range(base, base + self.sp))

    def cfg_ranks(self, rank: int) -> List[int]:
        shot, _, s = self.coords(rank)
        return [(shot * self.cfg + c) * self.sp + s for c in range(self.cfg)]

    def shots_of(self, rank: int, n_shots: int) -> List[int]:
        """Round-robin assignment of shot indices to this rank's shot group (batch_inference.py:45)."""
        return list(range(self.coords(rank)[0], n_shots, self.shots))

    @staticmethod
    def auto(world: int, n_shots: int = 1, cfg_on: bool = True, num_heads: int = 24) -> "Layout":
        """Prefer the axes that need no communication: shots first, then the CFG pair, Ulysses last."""
        shots = 1
        for cand in range(min(world, n_shots), 0, -1):
            if world % cand == 0:
                shots = cand
                break
        rest = world // shots
        cfg = 2 if (cfg_on and rest % 2 == 0) else 1
        sp = rest // cfg
        if num_heads % sp != 0:
            raise ValueError(f"sequence-parallel ways {sp} do not divide {num_heads} heads")
        return Layout(world, shots, cfg, sp)


class ParallelContext:
    """Creates the SP and CFG-pair process groups of a :class:`Layout` (every rank must construct it:
    ``dist.new_group`` is collective) and offers the two operations the denoise loop needs."""

    def __init__(self, layout: Layout, rank: Optional[int] = None, exchange: str = "p2p"):
        if not dist.is_initialized():
            raise RuntimeError("ParallelContext needs an initialised torch.distributed process group")
        if dist.get_world_size() != layout.world:
            raise ValueError(f"layout is for world {layout.world}, process group has {dist.get_world_size()}")
        self.layout = layout
        self.exchange = exchange
        self.rank = dist.get_rank() if rank is None else rank
        self.shot_group, self.cfg_rank, self.sp_rank = layout.coords(self.rank)
        self.sp_group = None
        self.cfg_group = None
        seen = set()
        for r in range(layout.world):  # identical creation order on every rank
            ranks = tuple(layout.sp_ranks(r))
            if ranks in seen:
                continue
            seen.add(ranks)
            if layout.sp > 1:
                g = dist.new_group(list(ranks))
                if self.rank in ranks:
                    self.sp_group = g
        seen.clear()
        for r in range(layout.world):
            ranks = tuple(layout.cfg_ranks(r))
            if ranks in seen:
                continue
            seen.add(ranks)
            if layout.cfg > 1:
                g = dist.new_group(list(ranks))
                if self.rank in ranks:
                    self.cfg_group = g
        self._pair_buf = None

    def sequence_parallel(self):
        """The Ulysses group of this rank as a ``fairygen_b200.sp.SequenceParallel`` (None when sp == 1)."""
        if self.sp_group is None:
            return None
        from .sp import SequenceParallel

        return SequenceParallel(self.sp_group, exchange=self.exchange)

    # ---- CFG pair ---------------------------------------------------------------------------------
    def forward_pair(self, engine, latents, timestep, context_pos, context_neg, fuse: bool):
        """Run THIS rank's half of the guidance pair and exchange predictions with the partner.
        Returns (noise_pos, noise_neg) on every rank — what PIPE:296-301 computes sequentially."""
        if self.cfg_group is None:
            npos = engine.forward(latents, timestep, context_pos, fuse)
            nneg = engine.forward(latents, timestep, context_neg, fuse) if context_neg is not None else None
            return npos, nneg
        if context_neg is None:
            raise ValueError("CFG-parallel forward_pair needs a negative context (cfg_scale != 1 with a CFG pair layout)")
        mine = engine.forward(latents, timestep, context_pos if self.cfg_rank == 0 else context_neg, fuse)
        mine = mine.contiguous()
        n = mine.shape[0]
        shape = (2 * n,) + tuple(mine.shape[1:])  # concatenated along dim 0: [positive; negative]
        if self._pair_buf is None or tuple(self._pair_buf.shape) != shape or self._pair_buf.dtype != mine.dtype:
            self._pair_buf = torch.empty(shape, dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(self._pair_buf, mine, group=self.cfg_group)
        return self._pair_buf[:n], self._pair_buf[n:]


def denoise_shots(ctx: ParallelContext, denoiser_factory, shots: Sequence[dict]) -> List[Tuple[int, torch.Tensor]]:
    """``batch_inference.py:45-56`` with the shots spread over the layout's shot groups.

    ``shots[i]`` = dict(latents=, context_pos=, context_neg=, first_frame_latents=); ``denoiser_factory()``
    returns a ``WanDenoiser`` bound to this rank's engine.  Returns [(shot index, final latents)] for the
    shots this rank's group rendered (every rank of the group holds the same replicated result)."""
    den = denoiser_factory()
    den.cfg_group = ctx if ctx.layout.cfg > 1 else None
    done = []
    for i in ctx.layout.shots_of(ctx.rank, len(shots)):
        s = shots[i]
        done.append((i, den(s["latents"], s["context_pos"], s.get("context_neg"), s.get("first_frame_latents"))))
    return done
