"""Stage-2 motion-LoRA fine-tune step of the DiT (BASELINE config 5) on the sm_100a kernels.

What the reference does (animation/diffsynth): ``DiffusionTrainingModule`` injects rank-32 LoRA adapters into the
300 Linears "q,k,v,o,ffn.0,ffn.2", loads the stage-1 (A1, B1), adds a zero-initialised trainable ``lora_B2`` per
Linear and monkey-patches the stage-2 forward (diffusion/training_module.py:266-352, TMOD); ``FlowMatchSFTLoss``
(diffusion/loss.py:5-21, LOSS) noises the latents, calls ``pipe.model_fn`` under autograd and back-propagates the
weighted MSE; each DiTBlock is re-computed in backward (pipelines/wan_video.py:1348-1360).

Here:
  * the frozen stage-1 adapter is folded once into W₁ = W + s·B1·A1 (``fgb_lora_merge``); the trainable term
    s·(B2*mask*2)(A1 x) rides along as ONE extra 64-wide K-block of the same tcgen05 GEMM (``fgb_gemm_bf16_ex``:
    second operand pair A2 = t = A1 x, W2 = masked B2), and in the backward as dX = dY·W₁ + (dY·B2eff)·A1
    (``fgb_gemm_dgrad_ex``) — no 10 GB weight re-merge per step, and the rank-32 side GEMM ``t = A1 x`` is the
    activation the B2 gradient needs anyway;
  * the forward keeps every block's activations (23 GB at 480x832x49; B200 has 180 GB), so nothing is re-computed
    unless ``recompute=True`` asks for the reference's checkpointing schedule;
  * the backward is hand-written: ``fgb_attn_bwd`` (tensor cores), ``fgb_gemm_dgrad``, fused LN / RMSNorm+RoPE /
    GELU / gate backward kernels, and ``fgb_lora_wgrad`` for dB2 = (dyᵀ A1x) * mask * 2 in fp32.
Only ``lora_B2`` receives gradients (TMOD:279-307); base weights, A1, B1, norms and modulations are frozen.

With an engine built on a ``SequenceParallel(exchange="p2p")`` group the trainer splits the tokens of one video over the
group ("SP ... reuses the same exchange in backward"): the forward exchange of the inference path, and in the
backward O, dO -> head owners, ``fgb_attn_bwd`` on the local heads, dq|dk|dv -> token owners (``fgb_sp_return_heads``), all
as NVLink peer stores from our kernels, then one sum all-reduce of this backward's LoRA gradients.

``Stage2Trainer.model_fn`` wraps forward/backward in a ``torch.autograd.Function`` so the reference's own
``FlowMatchSFTLoss`` / ``loss.backward()`` drive it unchanged; ``Stage2Trainer.step`` runs the whole step (noise,
forward, loss, backward) without autograd.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from . import ops
from .engine import WanDiTEngine
from .ops import BF16, EPI_BIAS, EPI_GATED_RESIDUAL, EPI_RESIDUAL
from .scheduler import FlowMatchScheduler


def lora_targets(num_layers: int) -> Iterable[str]:
    """Module names adapted by --lora_target_modules "q,k,v,o,ffn.0,ffn.2" (peft suffix match, TMOD:35-42)."""
    for i in range(num_layers):
        for a in ("self_attn", "cross_attn"):
            for p in "qkvo":
                yield f"blocks.{i}.{a}.{p}"
        yield f"blocks.{i}.ffn.0"
        yield f"blocks.{i}.ffn.2"


class _Saved:
    """Activations of one block kept for its backward."""
    __slots__ = ("x_in", "a1", "qk_pre", "qkv", "o", "lse", "x1", "a2", "cq_pre", "cq", "ck_pre", "ckv", "co", "lse_c", "x2", "a3",
                 "z1", "h", "t_qkv", "t_o", "t_cq", "t_ckv", "t_co", "t_1", "t_2")


class Stage2Trainer:
    """stage = 2 (default): FairyGen's motion stage — A1, B1 frozen, ``lora_B2`` trainable, weight dropout 0.5 (TMOD:266-352).
    stage = 1: the identity stage — ``lora_A`` and ``lora_B`` trainable, weight dropout 0.8 on B (TMOD:200-264); the same
    kernels, with ``self.b2`` holding B, no frozen adapter to merge, and dA = (dY·Beff)ᵀ·X from the factor u that the
    dgrad K-extension computes anyway."""

    def __init__(self, engine: WanDiTEngine, lora: Dict[str, torch.Tensor], rank: int = 32, lora_alpha: Optional[float] = None,
                 dropout_prob: Optional[float] = None, recompute: bool = False, dp_group=None, stage: int = 2):
        if stage not in (1, 2):
            raise ValueError(f"stage must be 1 or 2, got {stage}")
        self.stage = stage
        dropout_prob = (0.5 if stage == 2 else 0.8) if dropout_prob is None else dropout_prob
        if not engine.loaded:
            raise RuntimeError("Stage2Trainer needs an engine with loaded (frozen) base weights")
        self.sp = engine.sp   # Ulysses group: tokens of ONE video split over its ranks, forward and backward (see _sp_*)
        if self.sp is not None and getattr(self.sp, "exchange", "nccl") != "p2p":
            raise NotImplementedError("sequence-parallel training uses the peer-memory exchange: SequenceParallel(exchange='p2p')")
        self.engine = engine
        self.cfg = engine.cfg
        self.rank = rank
        self.scaling = float((lora_alpha if lora_alpha is not None else rank) / rank)   # peft: alpha / r (TMOD:33-34)
        self.dropout_prob = float(dropout_prob)
        self.mask_mul = 1.0 / (1.0 - self.dropout_prob)                                 # TMOD:343
        self.recompute = recompute
        self.dp_group = dp_group
        dev, cfg = engine.device, engine.cfg
        d, f = cfg.dim, cfg.ffn_dim
        self.targets: List[str] = list(lora_targets(cfg.num_layers))
        out_dim = lambda t: f if t.endswith("ffn.0") else d  # noqa: E731
        # ---- A (frozen in stage 2, trainable in stage 1): one flat bf16 buffer laid out per fused GEMM, so that the
        #      [3r, D] / [2r, D] operands of the fused q|k|v and cross k|v side GEMMs are views, not copies
        g = lambda k: lora[k].detach().to(device=dev, dtype=BF16).contiguous()  # noqa: E731
        in_dim = lambda t: f if t.endswith("ffn.2") else d  # noqa: E731
        self.a_flat = torch.empty(sum(rank * in_dim(t) for t in self.targets), dtype=BF16, device=dev)
        self.grad_a_flat = torch.zeros(self.a_flat.numel(), dtype=torch.float32, device=dev) if stage == 1 else None
        self.a1: Dict[str, torch.Tensor] = {}
        self.grad_a: Dict[str, torch.Tensor] = {}
        self.b1: Dict[str, torch.Tensor] = {}
        self.a1_qkv, self.a1_ckv = [], []
        off = 0
        for i in range(cfg.num_layers):      # lora_targets() order: self q,k,v,o, cross q,k,v,o, ffn.0, ffn.2
            p = f"blocks.{i}."
            block_start = off
            for t in [p + "self_attn." + c for c in "qkvo"] + [p + "cross_attn." + c for c in "qkvo"] + [p + "ffn.0", p + "ffn.2"]:
                n = rank * in_dim(t)
                self.a1[t] = self.a_flat[off:off + n].view(rank, in_dim(t))
                if stage == 1:
                    self.grad_a[t] = self.grad_a_flat[off:off + n].view(rank, in_dim(t))
                off += n
            self.a1_qkv.append(self.a_flat[block_start:block_start + 3 * rank * d].view(3 * rank, d))
            ck = block_start + 5 * rank * d   # after self q,k,v,o and cross q
            self.a1_ckv.append(self.a_flat[ck:ck + 2 * rank * d].view(2 * rank, d))
        for t in self.targets:
            a = g(f"{t}.lora_A.default.weight")
            if a.shape[0] != rank:
                raise ValueError(f"{t}: LoRA rank {a.shape[0]} != {rank}")
            self.a1[t].copy_(a)
            self.b1[t] = g(f"{t}.lora_B.default.weight")
        # ---- trainable B2: one flat bf16 parameter buffer, fp32 gradient / Adam moments ---------------------------
        sizes = [out_dim(t) * rank for t in self.targets]
        total = sum(sizes)
        self.b2_flat = torch.zeros(total, dtype=BF16, device=dev)                       # zero init (TMOD:296-298)
        self.grad_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_m = self.adam_v = None
        self.adam_step = 0
        self.b2: Dict[str, torch.Tensor] = {}
        self.grad: Dict[str, torch.Tensor] = {}
        self.mask: Dict[str, torch.Tensor] = {}
        self._mask_flat = torch.ones(total, dtype=torch.uint8, device=dev)
        off = 0
        for t, n in zip(self.targets, sizes):
            self.b2[t] = self.b2_flat[off:off + n].view(-1, rank)
            self.grad[t] = self.grad_flat[off:off + n].view(-1, rank)
            self.mask[t] = self._mask_flat[off:off + n].view(-1, rank)
            off += n
        # ---- W1 = W + s*B1*A1 (static; same layouts as the engine's packed base weights) and the per-step masked B2
        #      operands of the GEMM K-extension (block-diagonal for the fused q|k|v and cross k|v GEMMs)
        e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)  # noqa: E731
        z = lambda *s: torch.zeros(*s, dtype=BF16, device=dev)  # noqa: E731
        r = rank
        if stage == 2:
            self.eff = [dict(wqkv=e(3 * d, d), wo=e(d, d), cwq=e(d, d), cwkv=e(2 * d, d), cwo=e(d, d), w1=e(f, d), w2=e(d, f))
                        for _ in range(cfg.num_layers)]
        else:   # stage 1: no frozen adapter — the GEMMs read the engine's base weights directly
            self.eff = [dict(wqkv=b.wqkv, wo=b.wo, cwq=b.cwq, cwkv=b.cwkv, cwo=b.cwo, w1=b.w1, w2=b.w2) for b in engine.blocks]
        self.b2e = [dict(qkv=z(3 * d, 3 * r), o=z(d, r), cq=z(d, r), ckv=z(2 * d, 2 * r), co=z(d, r), f1=z(f, r), f2=z(d, r))
                    for _ in range(cfg.num_layers)]
        if stage == 2:
            self._merge_static()
        else:
            self.load_b2({t: self.b1[t] for t in self.targets})   # stage 1 trains B itself (self.b2 is the trainable matrix)
        self.scheduler = FlowMatchScheduler("Wan")
        self.scheduler.set_timesteps(1000, training=True)                               # train.py / LOSS:6-9
        self._saved: List[Optional[_Saved]] = []
        self._shape_key = None
        self._fwd_state = None
        self.loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        self.kernel_launches = 0
        self._mask_calls = 0
        self._b2_table = None

    # ------------------------------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------------------------------
    def load_b2(self, b2: Dict[str, torch.Tensor]) -> None:
        for t in self.targets:
            self.b2[t].copy_(b2[t].to(device=self.engine.device, dtype=BF16))

    def set_masks(self, masks: Optional[Dict[str, torch.Tensor]] = None, seed: Optional[int] = None) -> None:
        """Weight-dropout keep-masks of this step: injected (parity tests) or drawn on the device from `seed`."""
        if masks is not None:
            for t in self.targets:
                self.mask[t].copy_(masks[t].to(device=self.engine.device, dtype=torch.uint8))
        else:
            if seed is None:   # a fresh mask per call, as torch.rand_like gives the reference (TMOD:338-342)
                self._mask_calls += 1
                seed = 0x5EED + 7919 * self._mask_calls
            ops.bernoulli_mask(self._mask_flat, self.dropout_prob, seed)
            self.kernel_launches += 1

    def zero_grad(self) -> None:
        self.grad_flat.zero_()
        if self.grad_a_flat is not None:
            self.grad_a_flat.zero_()

    def _merge_static(self) -> None:
        """W1 = W + s * B1 * A1 for all 300 Linears (row slices of the fused QKV / cross-KV weights) — once."""
        d = self.cfg.dim
        for i, (b, eff) in enumerate(zip(self.engine.blocks, self.eff)):
            p = f"blocks.{i}."
            for j, proj in enumerate("qkv"):
                self._merge_one(b.wqkv[j * d:(j + 1) * d], eff["wqkv"][j * d:(j + 1) * d], p + "self_attn." + proj)
            self._merge_one(b.wo, eff["wo"], p + "self_attn.o")
            self._merge_one(b.cwq, eff["cwq"], p + "cross_attn.q")
            for j, proj in enumerate("kv"):
                self._merge_one(b.cwkv[j * d:(j + 1) * d], eff["cwkv"][j * d:(j + 1) * d], p + "cross_attn." + proj)
            self._merge_one(b.cwo, eff["cwo"], p + "cross_attn.o")
            self._merge_one(b.w1, eff["w1"], p + "ffn.0")
            self._merge_one(b.w2, eff["w2"], p + "ffn.2")

    def _merge_one(self, w, w_eff, name) -> None:
        ops.lora_merge(w, self.a1[name], self.b1[name], None, None, w_eff, 1.0, self.scaling)

    def _b2_eff_table(self) -> torch.Tensor:
        """(offset into b2_flat, rows, destination pointer, destination stride) of every adapted Linear: where its masked,
        scaled B2 goes inside the K-extension operands (diagonal blocks for the fused q|k|v and cross k|v GEMMs)."""
        d, r = self.cfg.dim, self.rank
        base = self.b2_flat.data_ptr()
        rows = []

        def add(name, out):
            src = self.b2[name]
            rows.append(((src.data_ptr() - base) // 2, src.shape[0], out.data_ptr(), out.stride(0)))

        for i, be in enumerate(self.b2e):
            p = f"blocks.{i}."
            for j, proj in enumerate("qkv"):
                add(p + "self_attn." + proj, be["qkv"][j * d:(j + 1) * d, j * r:(j + 1) * r])
            add(p + "self_attn.o", be["o"])
            add(p + "cross_attn.q", be["cq"])
            for j, proj in enumerate("kv"):
                add(p + "cross_attn." + proj, be["ckv"][j * d:(j + 1) * d, j * r:(j + 1) * r])
            add(p + "cross_attn.o", be["co"])
            add(p + "ffn.0", be["f1"])
            add(p + "ffn.2", be["f2"])
        return torch.tensor(rows, dtype=torch.int64, device=self.engine.device)

    def merge(self) -> None:
        """Per step: B2eff = s * bf16(bf16(B2 * mask) * 2) (TMOD:343-346) into the K-extension operands — one launch for
        all 10 x num_layers Linears."""
        if self._b2_table is None:
            self._b2_table = self._b2_eff_table()
        ops.lora_b2_eff_batched(self.b2_flat, self._mask_flat, self._b2_table, self.rank, self.mask_mul, self.scaling)
        self.kernel_launches += 1

    # ------------------------------------------------------------------------------------------------------------
    # buffers
    # ------------------------------------------------------------------------------------------------------------
    def _alloc_saved(self, rows: int, n_ctx: int) -> _Saved:
        cfg, dev = self.cfg, self.engine.device
        d, f, r, H = cfg.dim, cfg.ffn_dim, self.rank, cfg.num_heads
        P = self.sp.world if self.sp is not None else 1
        e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)  # noqa: E731
        s = _Saved()
        s.x_in, s.a1, s.o, s.x1, s.a2, s.cq_pre, s.cq, s.co, s.x2, s.a3 = (e(rows, d) for _ in range(10))
        # q|k|v after norm + RoPE: [rows, 3*H*128], or under SP the exchanged matrix [rows*P tokens, 3*(H/P)*128]
        s.qk_pre, s.qkv = e(rows, 2 * d), e(rows * P, 3 * d // P)
        s.ck_pre, s.ckv = e(n_ctx, d), e(n_ctx, 2 * d)
        s.z1, s.h = e(rows, f), e(rows, f)
        s.t_qkv, s.t_o, s.t_cq, s.t_ckv, s.t_co, s.t_1, s.t_2 = e(rows, 3 * r), e(rows, r), e(rows, r), e(n_ctx, 2 * r), e(rows, r), e(rows, r), e(rows, r)
        s.lse = torch.zeros(H // P, ops.stat_rows(rows * P), dtype=torch.float32, device=dev)
        s.lse_c = torch.zeros(H, ops.stat_rows(rows), dtype=torch.float32, device=dev)
        return s

    def _buffers(self, rows: int, n_ctx: int):
        key = (rows, n_ctx)
        if self._shape_key != key:
            cfg, dev = self.cfg, self.engine.device
            d, f, H = cfg.dim, cfg.ffn_dim, cfg.num_heads
            e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)  # noqa: E731
            n_sets = 1 if self.recompute else cfg.num_layers
            self._saved = [self._alloc_saved(rows, n_ctx) for _ in range(n_sets)]
            self._x_ckpt = [e(rows, d) for _ in range(cfg.num_layers)] if self.recompute else None
            self._scratch = dict(dx=e(rows, d), t1=e(rows, d), t2=e(rows, d), dqkv=e(rows, 3 * d), dh=e(rows, f), dckv=e(n_ctx, 2 * d),
                                 delta=torch.empty(H, ops.stat_rows(rows), dtype=torch.float32, device=dev), x=e(rows, d),
                                 x_final=e(rows, d), a_head=e(rows, d), d_hrow=e(rows, cfg.out_dim * 4), u=e(max(rows, n_ctx), 3 * self.rank))
            if self.sp is not None:
                P = self.sp.world
                z = lambda *s: torch.zeros(*s, dtype=BF16, device=dev)  # noqa: E731
                # dloc: dq|dk|dv of this rank's heads for all tokens; its padded key rows are never written and stay zero
                self._scratch.update(qkv_tok=e(rows, 3 * d), dloc=z(rows * P, 3 * d // P), d_hrow_full=z(rows * P, cfg.out_dim * 4),
                                     delta_sp=torch.empty(H // P, ops.stat_rows(rows * P), dtype=torch.float32, device=dev))
            self._shape_key = key
        return self._scratch

    def _saved_for(self, i: int) -> _Saved:
        return self._saved[0 if self.recompute else i]

    # ------------------------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------------------------
    def _block_forward(self, i: int, x: torch.Tensor, s: _Saved, st) -> None:
        """DiTBlock.forward (DIT:213-229) with the stage-2 LoRA folded into `eff`; x is updated in place."""
        eng, cfg = self.engine, self.cfg
        d, H = cfg.dim, cfg.num_heads
        b, w = eng.blocks[i], self.eff[i]
        m0, m1, n_first, grid, ctx_emb = st["m0"][i], st["m1"][i], st["n_first"], st["grid"], st["ctx_emb"]
        be = self.b2e[i]
        s.x_in.copy_(x)
        ops.ln_modulate(x, s.a1, cfg.eps, m0[0], m0[1], m1[0], m1[1], n_first)
        ops.gemm(s.a1, self.a1_qkv[i], None, s.t_qkv)
        if self.sp is None:
            ops.gemm(s.a1, w["wqkv"], b.bqkv, s.qkv, a2=s.t_qkv, w2=be["qkv"])
            s.qk_pre.copy_(s.qkv[:, :2 * d])
            ops.rmsnorm_rope(s.qkv[:, :d], cfg.eps, b.nq, eng.rope_tab, grid, 0)
            ops.rmsnorm_rope(s.qkv[:, d:2 * d], cfg.eps, b.nk, eng.rope_tab, grid, 0)
            ops.head_norm_max(s.qkv[:, d:2 * d], st["kmax2"], H)
            ops.attention(s.qkv[:, :d], s.qkv[:, d:2 * d], s.qkv[:, 2 * d:], s.o, H, lse=s.lse, kmax2=st["kmax2"])
        else:
            self._sp_attention_forward(i, s, st)
        ops.gemm(s.o, self.a1[f"blocks.{i}.self_attn.o"], None, s.t_o)
        ops.gemm(s.o, w["wo"], b.bo, x, EPI_GATED_RESIDUAL, m0[2], m1[2], n_first, a2=s.t_o, w2=be["o"])
        s.x1.copy_(x)
        ops.ln_affine(x, s.a2, cfg.eps, b.n3w, b.n3b)
        ops.gemm(s.a2, self.a1[f"blocks.{i}.cross_attn.q"], None, s.t_cq)
        ops.gemm(s.a2, w["cwq"], b.cbq, s.cq, a2=s.t_cq, w2=be["cq"])
        s.cq_pre.copy_(s.cq)
        ops.rmsnorm_rope(s.cq, cfg.eps, b.cnq)
        ops.gemm(ctx_emb, self.a1_ckv[i], None, s.t_ckv)
        ops.gemm(ctx_emb, w["cwkv"], b.cbkv, s.ckv, a2=s.t_ckv, w2=be["ckv"])
        s.ck_pre.copy_(s.ckv[:, :d])
        ops.rmsnorm_rope(s.ckv[:, :d], cfg.eps, b.cnk)
        ops.head_norm_max(s.ckv[:, :d], st["kmax2"], H)
        ops.attention(s.cq, s.ckv[:, :d], s.ckv[:, d:], s.co, H, lse=s.lse_c, kmax2=st["kmax2"])
        ops.gemm(s.co, self.a1[f"blocks.{i}.cross_attn.o"], None, s.t_co)
        ops.gemm(s.co, w["cwo"], b.cbo, x, EPI_RESIDUAL, a2=s.t_co, w2=be["co"])
        s.x2.copy_(x)
        ops.ln_modulate(x, s.a3, cfg.eps, m0[3], m0[4], m1[3], m1[4], n_first)
        ops.gemm(s.a3, self.a1[f"blocks.{i}.ffn.0"], None, s.t_1)
        ops.gemm(s.a3, w["w1"], b.b1, s.z1, EPI_BIAS, a2=s.t_1, w2=be["f1"])
        ops.gelu_tanh(s.z1, s.h)
        ops.gemm(s.h, self.a1[f"blocks.{i}.ffn.2"], None, s.t_2)
        ops.gemm(s.h, w["w2"], b.b2, x, EPI_GATED_RESIDUAL, m0[5], m1[5], n_first, a2=s.t_2, w2=be["f2"])
        self.kernel_launches += 34

    def _sp_attention_forward(self, i: int, s: _Saved, st) -> None:
        """Self-attention of block i under Ulysses SP (xdit_context_parallel.py:125-146): q, k leave for the rank that owns
        their head straight from the RMSNorm+RoPE kernel, v by a plain scatter; attention runs on H/P heads x all tokens and
        its epilogue stores each token row back to its owner.  Kept for the backward: the pre-norm q|k (token-major), the
        exchanged q|k|v (head-major), the log-sum-exp of the local heads and the gathered output."""
        eng, cfg, sp, sc = self.engine, self.cfg, self.sp, self._scratch
        d, H, P, rk = cfg.dim, cfg.num_heads, self.sp.world, self.sp.rank
        b, w, be = eng.blocks[i], self.eff[i], self.b2e[i]
        hpr, wloc, S, rows = H // P, d // P, st["S"], s.a1.shape[0]
        ar, qkv, dev = sp.arena, sc["qkv_tok"], eng.device
        ops.gemm(s.a1, w["wqkv"], b.bqkv, qkv, a2=s.t_qkv, w2=be["qkv"])
        s.qk_pre.copy_(qkv[:, :2 * d])
        ops.rmsnorm_rope_scatter(qkv[:, :d], cfg.eps, b.nq, eng.rope_tab, st["grid"], st["tok0"], ar.recv_ptrs, P, rk, 0, 3)
        ops.rmsnorm_rope_scatter(qkv[:, d:2 * d], cfg.eps, b.nk, eng.rope_tab, st["grid"], st["tok0"], ar.recv_ptrs, P, rk, 1, 3)
        ops.sp_scatter_heads(qkv[:, 2 * d:], ar.recv_ptrs, H, 1, P, rk, 2, 3)
        sp.barrier(0, dev)
        recv, kmax2 = ar.recv, st["kmax2"][:hpr]
        ops.head_norm_max(recv[:S, wloc:2 * wloc], kmax2, hpr)
        ops.attention_scatter(recv[:, :wloc], recv[:S, wloc:2 * wloc], recv[:S, 2 * wloc:], ar.o_ptrs, d, rows, rk * wloc, hpr,
                              kmax2=kmax2, lse=s.lse)
        s.qkv.copy_(recv)          # before the barrier: the peers' next scatter may overwrite recv right after it
        sp.barrier(1, dev)
        s.o.copy_(ar.o)            # complete once every peer's attention epilogue has passed the barrier
        self.kernel_launches += 2

    def _sp_attention_backward(self, i: int, s: _Saved, d_o: torch.Tensor, st) -> torch.Tensor:
        """Reverse exchange: O and dO go to the head owners (same scatter as v in the forward), fgb_attn_bwd runs on the local
        heads over all tokens, and dq|dk|dv return token-major into the owners' arena (fgb_sp_return_heads).  Returns the
        [rows, 3*D] gradient matrix of this rank's tokens (a view of the arena)."""
        cfg, sp, sc = self.cfg, self.sp, self._scratch
        d, H, P, rk = cfg.dim, cfg.num_heads, self.sp.world, self.sp.rank
        hpr, wloc, S, rows = H // P, d // P, st["S"], s.a1.shape[0]
        ar, dloc, dev = sp.arena, sc["dloc"], self.engine.device
        ops.sp_scatter_heads(s.o, ar.recv_ptrs, H, 1, P, rk, 0, 3)
        ops.sp_scatter_heads(d_o, ar.recv_ptrs, H, 1, P, rk, 1, 3)
        sp.barrier(0, dev)
        recv = ar.recv
        ops.attention_bwd(s.qkv[:, :wloc], s.qkv[:S, wloc:2 * wloc], s.qkv[:S, 2 * wloc:], recv[:, :wloc], recv[:, wloc:2 * wloc], s.lse,
                          dloc[:, :wloc], dloc[:S, wloc:2 * wloc], dloc[:S, 2 * wloc:], hpr, delta=sc["delta_sp"])
        ops.sp_return_heads(dloc, ar.dqkv_ptrs, 3 * d, rows, H, 3, P, rk)
        sp.barrier(1, dev)
        self.kernel_launches += 4
        return ar.dqkv

    def forward_train(self, latents: torch.Tensor, timestep: torch.Tensor, context: torch.Tensor,
                      fuse_vae_embedding_in_latents: bool = True) -> torch.Tensor:
        """model_fn_wan_video (PIPE:1217-1388) with saved activations; call set_masks() + merge() first."""
        eng, cfg = self.engine, self.cfg
        dev, d = eng.device, cfg.dim
        if latents.dim() != 5 or latents.shape[0] != 1 or latents.shape[1] != cfg.in_dim:
            raise ValueError(f"latents must be (1,{cfg.in_dim},F,H,W), got {tuple(latents.shape)}")
        f, h, w = latents.shape[2], latents.shape[3] // 2, latents.shape[4] // 2
        grid, S = (f, h, w), f * h * w
        sp = self.sp
        P, rk = (sp.world, sp.rank) if sp is not None else (1, 0)
        rows = -(-S // P)              # tokens of this rank (contiguous split, zero-padded tail: PIPE:1312-1315)
        tok0 = rk * rows
        if sp is not None:             # the arena with the backward region must exist before the engine maps its workspace
            old = sp.arena
            if sp.peer_arena(rows, cfg.num_heads, dev, backward=True) is not old:
                eng._ws = {}
        ws = eng._workspace(rows, rows * P)
        per_token = bool(cfg.seperated_timestep and fuse_vae_embedding_in_latents)
        R = 2 if per_token else 1
        mod_tab, head_tab = eng._time_tables(ws, timestep, R)
        r_main = R - 1
        n_first = max(0, min(rows, h * w - tok0)) if per_token else 0
        ctx_emb = eng.text_embedding(context)
        sc = self._buffers(rows, ctx_emb.shape[0])
        L = cfg.num_layers
        st = dict(grid=grid, S=S, tok0=tok0, n_first=n_first, ctx_emb=ctx_emb, kmax2=ws["kmax2"],
                  m0=[mod_tab[0, i].view(6, d) for i in range(L)], m1=[mod_tab[r_main, i].view(6, d) for i in range(L)],
                  h0=head_tab[0].view(2, d), h1=head_tab[r_main].view(2, d), lat_dtype=latents.dtype)
        x = sc["x"]
        lat = latents[0].to(device=dev, dtype=BF16).contiguous()
        ops.patchify_rows(lat, ws["prow"], grid, tok0)
        ops.gemm(ws["prow"], eng.w_patch, eng.b_patch, x)
        for i in range(L):
            if self.recompute:
                self._x_ckpt[i].copy_(x)
            self._block_forward(i, x, self._saved_for(i), st)
        sc["x_final"].copy_(x)
        ops.ln_modulate(x, sc["a_head"], cfg.eps, st["h0"][0], st["h0"][1], st["h1"][0], st["h1"][1], n_first)
        ops.gemm(sc["a_head"], eng.w_head, eng.b_head, ws["hrow"])
        hrow = ws["hrow"] if sp is None else sp.all_gather_rows(ws["hrow"], ws["hgather"])   # PIPE:1379-1382
        out = torch.empty(cfg.out_dim, f, 2 * h, 2 * w, dtype=BF16, device=dev)
        ops.unpatchify(hrow, out, grid)
        self.kernel_launches += 5
        self._fwd_state = st
        return out.unsqueeze(0)

    # ------------------------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------------------------
    def _dgrad(self, dy, w1, b2e, a1, dx, x_in=None, names=()) -> None:
        """dX = dY·W1 + (dY·B2eff)·A1: the rank-r factor u = dY·B2eff first (one narrow dgrad), then the main dgrad with u·A1
        folded in as an extra K-block (dx = None: only u is needed). Stage 1: dA_j += u_jᵀ·X for the Linears in `names`."""
        u = self._scratch["u"][:dy.shape[0], :b2e.shape[1]]
        ops.gemm_dgrad(dy, b2e, u)
        if dx is not None:
            ops.gemm_dgrad(dy, w1, dx, u=u, a1=a1)
        self.kernel_launches += 2
        if self.stage == 1:
            r = self.rank
            for j, name in enumerate(names):
                ops.lora_wgrad(x_in, u[:, j * r:(j + 1) * r], self.grad_a[name], None, 1.0, transpose=True)
                self.kernel_launches += 1

    def _wgrad(self, dy, t, name) -> None:
        ops.lora_wgrad(dy, t, self.grad[name], self.mask[name], self.mask_mul * self.scaling)
        self.kernel_launches += 1

    def _block_backward(self, i: int, s: _Saved, st) -> None:
        """dx (grad w.r.t. the block output) -> dx (grad w.r.t. the block input), accumulating dB2 of its 10 Linears."""
        eng, cfg, sc = self.engine, self.cfg, self._scratch
        d, H, r = cfg.dim, cfg.num_heads, self.rank
        b, w = eng.blocks[i], self.eff[i]
        m0, m1, n_first, grid = st["m0"][i], st["m1"][i], st["n_first"], st["grid"]
        dx, t1, t2, dqkv, dh, dckv, delta = sc["dx"], sc["t1"], sc["t2"], sc["dqkv"], sc["dh"], sc["dckv"], sc["delta"]
        p = f"blocks.{i}."
        be = self.b2e[i]
        # ---- feed-forward branch: x3 = x2 + gate_mlp * ffn2(gelu(ffn0(a3)))                       (DIT:227-228)
        ops.mul_gate(dx, t1, m0[5], m1[5], n_first)
        self._wgrad(t1, s.t_2, p + "ffn.2")
        self._dgrad(t1, w["w2"], be["f2"], self.a1[p + "ffn.2"], dh, s.h, (p + "ffn.2",))
        ops.gelu_tanh_bwd(s.z1, dh, dh)
        self._wgrad(dh, s.t_1, p + "ffn.0")
        self._dgrad(dh, w["w1"], be["f1"], self.a1[p + "ffn.0"], t1, s.a3, (p + "ffn.0",))
        ops.ln_bwd(s.x2, t1, dx, cfg.eps, m0[4], m1[4], n_first, affine=False, dres=dx)
        # ---- cross-attention branch: x2 = x1 + o(attn(norm_q(q(a2)), norm_k(k(ctx)), v(ctx)))      (DIT:226)
        self._wgrad(dx, s.t_co, p + "cross_attn.o")
        self._dgrad(dx, w["cwo"], be["co"], self.a1[p + "cross_attn.o"], t1, s.co, (p + "cross_attn.o",))
        ops.attention_bwd(s.cq, s.ckv[:, :d], s.ckv[:, d:], s.co, t1, s.lse_c, t2, dckv[:, :d], dckv[:, d:], H, delta=delta)
        ops.rmsnorm_rope_bwd(s.cq_pre, t2, cfg.eps, b.cnq)
        ops.rmsnorm_rope_bwd(s.ck_pre, dckv[:, :d], cfg.eps, b.cnk)
        self._wgrad(t2, s.t_cq, p + "cross_attn.q")
        self._wgrad(dckv[:, :d], s.t_ckv[:, :r], p + "cross_attn.k")
        self._wgrad(dckv[:, d:], s.t_ckv[:, r:], p + "cross_attn.v")
        if self.stage == 1:   # the context has no gradient, but A of cross k / v does: only the factor u is needed
            self._dgrad(dckv, None, be["ckv"], None, None, st["ctx_emb"], (p + "cross_attn.k", p + "cross_attn.v"))
        self._dgrad(t2, w["cwq"], be["cq"], self.a1[p + "cross_attn.q"], t1, s.a2, (p + "cross_attn.q",))
        ops.ln_bwd(s.x1, t1, dx, cfg.eps, b.n3w, None, 0, affine=True, dres=dx)
        # ---- self-attention branch: x1 = x0 + gate_msa * o(attn(rope(norm_q(q(a1))), ...))        (DIT:224-225)
        ops.mul_gate(dx, t1, m0[2], m1[2], n_first)
        self._wgrad(t1, s.t_o, p + "self_attn.o")
        self._dgrad(t1, w["wo"], be["o"], self.a1[p + "self_attn.o"], t2, s.o, (p + "self_attn.o",))
        if self.sp is None:
            ops.attention_bwd(s.qkv[:, :d], s.qkv[:, d:2 * d], s.qkv[:, 2 * d:], s.o, t2, s.lse, dqkv[:, :d], dqkv[:, d:2 * d],
                              dqkv[:, 2 * d:], H, delta=delta)
        else:
            dqkv = self._sp_attention_backward(i, s, t2, st)
        ops.rmsnorm_rope_bwd(s.qk_pre[:, :d], dqkv[:, :d], cfg.eps, b.nq, eng.rope_tab, grid, st["tok0"])
        ops.rmsnorm_rope_bwd(s.qk_pre[:, d:], dqkv[:, d:2 * d], cfg.eps, b.nk, eng.rope_tab, grid, st["tok0"])
        for j, proj in enumerate("qkv"):
            self._wgrad(dqkv[:, j * d:(j + 1) * d], s.t_qkv[:, j * r:(j + 1) * r], p + "self_attn." + proj)
        self._dgrad(dqkv, w["wqkv"], be["qkv"], self.a1_qkv[i], t1, s.a1, tuple(p + "self_attn." + c for c in "qkv"))
        ops.ln_bwd(s.x_in, t1, dx, cfg.eps, m0[1], m1[1], n_first, affine=False, dres=dx)
        self.kernel_launches += 20

    def backward_train(self, dpred: torch.Tensor) -> None:
        """Back-propagate dL/dprediction (1,C,F,H,W) through head and blocks; adds into ``self.grad`` (fp32)."""
        st = self._fwd_state
        if st is None:
            raise RuntimeError("backward_train called before forward_train")
        eng, cfg, sc = self.engine, self.cfg, self._scratch
        grid, n_first = st["grid"], st["n_first"]
        dp = dpred.reshape(cfg.out_dim, grid[0], 2 * grid[1], 2 * grid[2]).to(dtype=BF16).contiguous()
        if self.sp is None:
            d_hrow = ops.unpatchify_bwd(dp, sc["d_hrow"], grid)
        else:   # every rank holds the whole prediction gradient; it keeps the rows of its own tokens (padded rows stay zero)
            rows = sc["d_hrow"].shape[0]
            d_hrow = ops.unpatchify_bwd(dp, sc["d_hrow_full"], grid)[st["tok0"]:st["tok0"] + rows]
        ops.gemm_dgrad(d_hrow, eng.w_head, sc["t1"])
        ops.ln_bwd(sc["x_final"], sc["t1"], sc["dx"], cfg.eps, st["h0"][1], st["h1"][1], n_first, affine=False, dres=None)
        self.kernel_launches += 3
        # group reductions act on THIS backward's gradients only: what earlier micro-steps accumulated is set aside
        reduce_groups = self.sp is not None or self.dp_group is not None
        flats = [g for g in (self.grad_flat, self.grad_a_flat) if g is not None]
        held = [g.clone() for g in flats] if reduce_groups else []
        if reduce_groups:
            for g in flats:
                g.zero_()
        for i in reversed(range(cfg.num_layers)):
            s = self._saved_for(i)
            if self.recompute:   # the reference's per-block checkpointing (PIPE:1348-1360): rebuild the activations
                sc["x"].copy_(self._x_ckpt[i])
                self._block_forward(i, sc["x"], s, st)
            self._block_backward(i, s, st)
        self._fwd_state = None
        if reduce_groups:
            import torch.distributed as dist

            for gflat, prev in zip(flats, held):
                if self.sp is not None:   # each rank saw only its tokens (and, for cross k / v, its queries): the gradients add up
                    dist.all_reduce(gflat, group=self.sp.group)
                if self.dp_group is not None:
                    dist.all_reduce(gflat, group=self.dp_group)
                    gflat.div_(dist.get_world_size(self.dp_group))   # DDP averages (accelerate, train.py)
                gflat.add_(prev)

    # ------------------------------------------------------------------------------------------------------------
    # whole step without autograd (LOSS:5-21)
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, input_latents: torch.Tensor, noise: torch.Tensor, timestep_id: int, context: torch.Tensor,
             masks: Optional[Dict[str, torch.Tensor]] = None, seed: Optional[int] = None,
             fuse_vae_embedding_in_latents: bool = True, return_pred: bool = False):
        """One forward + backward of FlowMatchSFTLoss with the random draws (timestep id, noise, masks) given."""
        dev = self.engine.device
        x0 = input_latents.to(device=dev, dtype=BF16).contiguous()
        nz = noise.to(device=dev, dtype=BF16).contiguous()
        sched = self.scheduler
        # LOSS:9 casts the sampled timestep to the pipeline dtype (bf16); add_noise / training_weight then look the
        # schedule index up again by argmin (FM:164-179), which can land on a neighbouring entry — reproduced as is
        t_bf16 = sched.timesteps[timestep_id:timestep_id + 1].to(BF16)
        timestep = t_bf16.to(torch.float32)
        index = sched._index(t_bf16)
        sigma = float(sched.sigmas[index])
        weight = float(sched.linear_timesteps_weights[index])
        latents, target = torch.empty_like(x0), torch.empty_like(x0)
        ops.fm_noise_target(x0, nz, sigma, latents, target)
        self.set_masks(masks, seed)
        self.merge()
        pred = self.forward_train(latents, timestep, context, fuse_vae_embedding_in_latents)
        dpred = torch.empty_like(pred)
        ops.mse_loss_grad(pred, target, weight, self.loss_buf, dpred)
        self.kernel_launches += 2
        self.backward_train(dpred)
        return (self.loss_buf, pred) if return_pred else self.loss_buf

    def optimizer_step(self, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2) -> None:
        if self.adam_m is None:
            self.adam_m, self.adam_v = torch.zeros_like(self.grad_flat), torch.zeros_like(self.grad_flat)
        self.adam_step += 1
        ops.adamw_step(self.b2_flat, self.grad_flat, self.adam_m, self.adam_v, lr, betas[0], betas[1], eps, weight_decay, self.adam_step)
        self.kernel_launches += 1
        if self.stage == 1:
            if getattr(self, "adam_ma", None) is None:
                self.adam_ma, self.adam_va = torch.zeros_like(self.grad_a_flat), torch.zeros_like(self.grad_a_flat)
            ops.adamw_step(self.a_flat, self.grad_a_flat, self.adam_ma, self.adam_va, lr, betas[0], betas[1], eps, weight_decay,
                           self.adam_step)
            self.kernel_launches += 1

    # ------------------------------------------------------------------------------------------------------------
    # autograd bridge: pipe.model_fn under torch.enable_grad (LOSS:17)
    # ------------------------------------------------------------------------------------------------------------
    def model_fn(self, b2_params: List[torch.nn.Parameter], latents, timestep, context, fuse_vae_embedding_in_latents=True,
                 masks=None, seed=None):
        """Prediction with a grad_fn: ``loss.backward()`` fills ``p.grad`` of the given B2 parameters (ordered as
        ``self.targets``), so the reference's FlowMatchSFTLoss + optimizer loop runs unchanged."""
        return _Stage2Function.apply(self, latents, timestep, context, fuse_vae_embedding_in_latents, masks, seed, *b2_params)


class _Stage2Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trainer: Stage2Trainer, latents, timestep, context, fuse, masks, seed, *b2_params):
        for t, p in zip(trainer.targets, b2_params):
            trainer.b2[t].copy_(p.detach().to(BF16))
        trainer.set_masks(masks, seed)
        trainer.merge()
        ctx.trainer = trainer
        ctx.n = len(b2_params)
        ctx.dtypes = [p.dtype for p in b2_params]
        return trainer.forward_train(latents, timestep.to(torch.float32), context, fuse).to(latents.dtype)

    @staticmethod
    def backward(ctx, dpred):
        tr = ctx.trainer
        tr.zero_grad()
        with torch.no_grad():
            tr.backward_train(dpred)
        grads = tuple(tr.grad[t].to(dt) for t, dt in zip(tr.targets, ctx.dtypes))
        return (None, None, None, None, None, None, None) + grads
