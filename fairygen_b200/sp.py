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


class PeerArena:
    """Caller-owned exchange arena of one rank, mapped into every peer of the Ulysses group through CUDA IPC:

        [ recv : s_pad x 3*(H/P)*128 bf16 | o : rows x H*128 bf16 | flags : 2 x 64 int32 ]

    ``recv`` is where the peers' fgb_sp_scatter_heads stores land (it is directly the q|k|v matrix the attention
    kernel reads), ``o`` is where the peers' attention epilogues store this rank's token rows, ``flags`` carries the
    epochs of fgb_sp_barrier.  torch.distributed is used once, to swap the 64-byte IPC handles."""

    def __init__(self, group, world: int, rank: int, rows: int, heads: int, device, backward: bool = False):
        self.group, self.world, self.rank, self.rows, self.device = group, world, rank, rows, device
        self.backward = backward
        s_pad = rows * world
        wloc = 3 * (heads // world) * 128
        d = heads * 128
        self.off_recv = 0
        self.off_o = (s_pad * wloc * 2 + 255) // 256 * 256
        self.off_dqkv = self.off_o + (rows * d * 2 + 255) // 256 * 256      # training only: [rows, 3*H*128] gradient matrix
        self.off_stats = self.off_dqkv + ((rows * 3 * d * 2 + 255) // 256 * 256 if backward else 0)   # fp32 [2][s_pad]: full-row sums of squares of q, k
        self.off_flags = self.off_stats + (2 * s_pad * 4 + 255) // 256 * 256
        total = self.off_flags + 2 * 64 * 4 + 64
        self.buf = self._allocate(total, device)
        self.recv = self.buf[self.off_recv:self.off_recv + s_pad * wloc * 2].view(torch.bfloat16).view(s_pad, wloc)
        self.o = self.buf[self.off_o:self.off_o + rows * d * 2].view(torch.bfloat16).view(rows, d)
        self.dqkv = self.buf[self.off_dqkv:self.off_dqkv + rows * 3 * d * 2].view(torch.bfloat16).view(rows, 3 * d) if backward else None
        self.stats = self.buf[self.off_stats:self.off_stats + 2 * s_pad * 4].view(torch.float32).view(2, s_pad)
        self.status = self.buf[self.off_flags + 2 * 64 * 4:self.off_flags + 2 * 64 * 4 + 4].view(torch.int32)   # barrier time-out report
        self.rowsq = torch.zeros(2, rows, dtype=torch.float32, device=device)   # local: filled by the q|k|v GEMM epilogue
        handle, offset = ops.ipc_export(self.buf)
        everyone = [None] * world
        dist.all_gather_object(everyone, (handle, offset), group=group)
        self._opened = []
        bases = []
        for q, (h, off) in enumerate(everyone):
            if q == rank:
                bases.append(self.buf.data_ptr())
            else:
                ptr = ops.ipc_open(device, h, off)
                self._opened.append((ptr, off))
                bases.append(ptr)
        self.recv_ptrs = [b + self.off_recv for b in bases]
        self.o_ptrs = [b + self.off_o for b in bases]
        self.dqkv_ptrs = [b + self.off_dqkv for b in bases]
        self.stats_ptrs = [b + self.off_stats for b in bases]
        self.flag_ptrs = [[b + self.off_flags + which * 64 * 4 for b in bases] for which in (0, 1)]
        self.epoch = 0
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)
        dist.barrier(group=group)   # every arena is zeroed and mapped before the first peer store

    @staticmethod
    def _allocate(nbytes: int, device) -> torch.Tensor:
        """Backing store of the arena: zeroed device memory (a hook: tests/test_engine_p2p_host.py maps it into shared host
        memory so that the peer-store exchange can run between CPU processes)."""
        return torch.zeros(nbytes, dtype=torch.uint8, device=device)

    def close(self):
        for ptr, off in self._opened:
            ops.ipc_close(self.device, ptr, off)
        self._opened = []


class SequenceParallel:
    def __init__(self, group=None, exchange: str = "p2p", overlap: bool = True):
        """exchange = "p2p": NVLink peer stores issued by our own kernels (scatter + fused attention epilogue);
        "nccl": pack / all-to-all / unpack (the plain-library variant, kept for comparison and for the CPU tests)."""
        if not dist.is_initialized():
            raise RuntimeError("SequenceParallel needs an initialised torch.distributed process group")
        if exchange not in ("p2p", "nccl"):
            raise ValueError(f"exchange must be 'p2p' or 'nccl', got {exchange!r}")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.exchange = exchange
        import os
        # p2p: run the QKV projection in two row chunks and scatter the first under the second (FGB_SP_OVERLAP=0: A/B switch)
        self.overlap = overlap and os.environ.get("FGB_SP_OVERLAP", "1") != "0"
        self.fused_send = os.environ.get("FGB_SP_FUSED", "1") != "0"   # q|k|v projection sends from its epilogue (A/B switch)
        self.arena = None
        self._side = None      # side stream + events of the overlapped scatter (created on first use)
        self._ev = None

    def peer_arena(self, rows: int, heads: int, device, backward: bool = False) -> PeerArena:
        if self.arena is None or self.arena.rows != rows or (backward and not self.arena.backward):
            if self.arena is not None:
                self.arena.close()
            self.arena = PeerArena(self.group, self.world, self.rank, rows, heads, device, backward=backward)
        return self.arena

    def barrier(self, which: int, device) -> None:
        """System-scope barrier of the exchange (which = 0 / 1: the two alternating flag sets)."""
        ar = self.arena
        if which == 0:
            ar.epoch += 1
        ops.sp_barrier(device, ar.flag_ptrs[which], self.world, self.rank, ar.epoch, ar.status)

    def check(self) -> None:
        """Raise if a barrier of the exchange gave up on a peer (fgb_sp_barrier_status / fgb_sp_stats_barrier write the epoch
        of the first time-out into the arena's status word and let the stream run on): everything computed since then read
        stale peer data.  One 4-byte device read — call it where a result leaves the GPU (WanDenoiser does, after the loop)."""
        ar = self.arena
        if ar is None or self.exchange != "p2p":
            return
        epoch = int(ar.status.item())
        if epoch != 0:
            raise RuntimeError(f"fairygen_b200 sequence parallel: rank {self.rank} of {self.world} waited ~30 s at exchange barrier "
                               f"epoch {epoch} (now at {ar.epoch}) for a peer that never arrived — a rank died or fell out of step; "
                               "results since that epoch are invalid")

    # ---- collectives (thin: NCCL on device tensors, gloo in the CPU tests) ----------------------
    def all_to_all(self, recv: torch.Tensor, send: torch.Tensor) -> torch.Tensor:
        dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def all_gather_rows(self, rows: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        dist.all_gather_into_tensor(out, rows.contiguous(), group=self.group)
        return out

    # ---- the self-attention exchange ---------------------------------------------------------------
    def _scatter_rows(self, engine, qkv, r0: int, r1: int, norm) -> None:
        """Normalise / rotate / send rows [r0, r1) of this rank's fused q|k|v matrix to the head owners (p2p exchange).
        The kernels address the destination row as rank*rows_given + r; for a row sub-range the peer base pointers are
        advanced so that it lands on rank*rows + r0 + r of the receive matrix."""
        ar = self.arena
        heads = engine.cfg.num_heads
        d = heads * 128
        rows = qkv.shape[0]
        n = r1 - r0
        row_bytes = 3 * (heads // self.world) * 128 * 2
        shift = (self.rank * (rows - n) + r0) * row_bytes
        ptrs = ar.recv_ptrs if shift == 0 else [p + shift for p in ar.recv_ptrs]
        x = qkv[r0:r1]
        k = engine._k
        if norm is None:
            k("sp_scatter", ops.sp_scatter_heads, x, ptrs, heads, 3, self.world, self.rank)
            return
        eps, wq, wk, rope_tab, grid, tok0 = norm
        # q and k leave for their owners straight from the RMSNorm+RoPE registers; v needs a plain scatter
        k("rmsnorm_rope", ops.rmsnorm_rope_scatter, x[:, :d], eps, wq, rope_tab, grid, tok0 + r0, ptrs, self.world, self.rank, 0, 3)
        k("rmsnorm_rope", ops.rmsnorm_rope_scatter, x[:, d:2 * d], eps, wk, rope_tab, grid, tok0 + r0, ptrs, self.world, self.rank, 1, 3)
        k("sp_scatter", ops.sp_scatter_heads, x[:, 2 * d:], ptrs, heads, 1, self.world, self.rank, 2, 3)

    def attention(self, engine, ws, qkv: torch.Tensor, o: torch.Tensor, tokens: int, norm=None, qkv_gemm=None, fused=None) -> None:
        """qkv [rows, 3*H*128] -> o [rows, H*128].  norm = None: q, k are already RMS-normed + rotated.
        norm = (eps, weight_q, weight_k, rope_tab, grid, token_offset) (p2p only): the pre-norm q, k are normalised,
        rotated and sent by one fused kernel each.
        qkv_gemm (p2p only): callable (r0, r1) that launches the fused QKV projection for rows [r0, r1) of `qkv`.  The
        projection then runs in two row chunks and the NVLink scatter of the first chunk (a side stream) overlaps the
        tensor-core work of the second — the exchange is link-bound (~0.5 TB/s of peer stores), the projection is not.
        fused (p2p only, the default of the engine): (a, w_qkv, b_qkv, eps, weight_q, weight_k, rope_tab, grid) — the projection
        itself sends: see the comments below and fgb_gemm_qkv_scatter."""
        heads = engine.cfg.num_heads
        hpr = heads // self.world
        wloc = hpr * 128
        k = engine._k
        if self.exchange == "p2p":
            ar = self.arena            # created by the engine's workspace; ws["recv"] / ws["o"] are views of it
            rows = qkv.shape[0]
            if fused is not None:
                # send side fused into the projection: the GEMM epilogue TMA-stores every block into the head owner's receive
                # matrix; the full-row statistics of q and k travel with barrier 0; RMSNorm + RoPE happen on the receiver
                a, w, bias, eps, wq, wk, rope_tab, grid = fused
                k("gemm_qkv", ops.gemm_qkv_scatter, a, w, bias, heads * 128, ar.recv_ptrs, self.world, self.rank, ar.rowsq, sk_ws=ws.get("sk"))
                ar.epoch += 1
                kmax2 = ws["kmax2"][:hpr]
                k("sp_barrier", ops.sp_stats_barrier, qkv.device, ar.flag_ptrs[0], ar.stats_ptrs, ar.rowsq, rows, rows * self.world, kmax2,
                  hpr, self.world, self.rank, ar.epoch, ar.status)
                h0 = self.rank * wloc
                qmax2 = ws["qmax2"][0][:hpr] if ("qmax2" in ws and getattr(engine, "query_bounds", 0) & 1) else None
                k("rmsnorm_rope", ops.recv_norm_rope, ar.recv, tokens, hpr, ar.stats, heads * 128, eps, wq[h0:h0 + wloc], wk[h0:h0 + wloc],
                  rope_tab, grid, kmax2, qmax2)
                recv = ar.recv
                k("attn_self", ops.attention_scatter, recv[:, :wloc], recv[:tokens, wloc:2 * wloc], recv[:tokens, 2 * wloc:], ar.o_ptrs,
                  heads * 128, rows, self.rank * wloc, hpr, kmax2=kmax2, qmax2=qmax2)
                k("sp_barrier", ops.sp_barrier, qkv.device, ar.flag_ptrs[1], self.world, self.rank, ar.epoch, ar.status)
                return
            if qkv_gemm is None:
                self._scatter_rows(engine, qkv, 0, rows, norm)
            elif rows < 1024 or not self.overlap:
                qkv_gemm(0, rows)
                self._scatter_rows(engine, qkv, 0, rows, norm)
            else:
                half = ((rows // 2 + 255) // 256) * 256      # whole 256-row GEMM tiles in the first chunk
                main = torch.cuda.current_stream(qkv.device)
                if self._side is None:
                    self._side = torch.cuda.Stream(device=qkv.device)
                    self._ev = [torch.cuda.Event(), torch.cuda.Event()]
                qkv_gemm(0, half)
                self._ev[0].record(main)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(self._ev[0])
                    self._scatter_rows(engine, qkv, 0, half, norm)
                    self._ev[1].record(self._side)
                qkv_gemm(half, rows)
                self._scatter_rows(engine, qkv, half, rows, norm)
                main.wait_event(self._ev[1])
            ar.epoch += 1
            k("sp_barrier", ops.sp_barrier, qkv.device, ar.flag_ptrs[0], self.world, self.rank, ar.epoch, ar.status)
            recv = ar.recv
            kmax2 = ws["kmax2"][:hpr]
            k("head_norm_max", ops.head_norm_max, recv[:tokens, wloc:2 * wloc], kmax2, hpr)
            k("attn_self", ops.attention_scatter, recv[:, :wloc], recv[:tokens, wloc:2 * wloc], recv[:tokens, 2 * wloc:], ar.o_ptrs,
              heads * 128, rows, self.rank * wloc, hpr, kmax2=kmax2)
            k("sp_barrier", ops.sp_barrier, qkv.device, ar.flag_ptrs[1], self.world, self.rank, ar.epoch, ar.status)
            return
        if qkv_gemm is not None:
            qkv_gemm(0, qkv.shape[0])
        k("sp_pack", ops.sp_pack_heads, qkv, ws["send"], heads, 3, self.world)
        recv = k("sp_all_to_all", self.all_to_all, ws["recv"], ws["send"])     # [s_pad tokens, (q|k|v) x hpr x 128]
        kmax2 = ws["kmax2"][:hpr] if "kmax2" in ws else None   # bounded-score softmax, as on the single-GPU and p2p paths
        if kmax2 is not None:
            k("head_norm_max", ops.head_norm_max, recv[:tokens, wloc:2 * wloc], kmax2, hpr)
        k("attn_self", ops.attention, recv[:, :wloc], recv[:tokens, wloc:2 * wloc], recv[:tokens, 2 * wloc:], ws["o_full"], hpr, kmax2=kmax2)
        o_recv = k("sp_all_to_all", self.all_to_all, ws["o_recv"], ws["o_full"])  # [world][rows][hpr*128]
        k("sp_unpack", ops.sp_unpack_heads, o_recv, o, heads, 1, self.world)
