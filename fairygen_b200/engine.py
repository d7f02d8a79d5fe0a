"""WanDiTEngine — the DiT forward of the hot path, executed entirely by the sm_100a kernels.

Host-side mirror of ``model_fn_wan_video`` (reference animation/diffsynth/pipelines/wan_video.py
:1217-1388, TI2V branches) driving ``DiTBlock`` / ``Head`` (models/wan_video_dit.py:195-268).  The
engine owns packed bf16 weights (fused QKV, fused cross K|V) and per-shape workspaces (PyTorch is
the allocator); every FLOP and byte on the path goes through ``fairygen_b200.ops`` -> C ABI.

What is restructured relative to the reference (results identical up to bf16 rounding):
  * the per-token time MLP runs on the 2 distinct rows (t=0 for the first latent frame, t for the
    rest; PIPE:1218-1228) and every modulate/gate picks its row by token index — the (1,S,6,D)
    t_mod tensor (1 GB at S=27 280) is never built;
  * text_embedding and the cross-attention K/V projections are step-invariant and cached per
    context tensor (PIPE:1236; DIT:177-178);
  * RoPE is applied from a float32 (cos,sin) table inside the q/k RMSNorm kernel (DIT:91-96).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import ops
from .config import WanDiTConfig
from .ops import BF16, EPI_BIAS, EPI_BIAS_GELU_TANH, EPI_GATED_RESIDUAL, EPI_RESIDUAL


class _Block:
    __slots__ = ("wqkv", "bqkv", "wo", "bo", "nq", "nk", "cwq", "cbq", "cwkv", "cbkv", "cwo", "cbo", "cnq", "cnk",
                 "n3w", "n3b", "w1", "b1", "w2", "b2", "private")

    def own(self, attr: str) -> torch.Tensor:
        """The weight `attr` as a tensor this engine may write: load_state_dict() does not copy a weight that already is bf16 on
        the device (no second 10 GB), so wo / cwq / cwo / w1 / w2 may alias the caller's module — an in-place LoRA fusion into
        the engine must not reach the module's parameters (copy on first write; the concatenated wqkv / cwkv are always ours)."""
        if attr not in self.private:
            setattr(self, attr, getattr(self, attr).clone())
            self.private.add(attr)
        return getattr(self, attr)


def _dev_bf16(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=BF16).contiguous()


class WanDiTEngine:
    def __init__(self, cfg: WanDiTConfig, device="cuda", sp=None):
        cfg.validate()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("WanDiTEngine needs a CUDA (sm_100a) device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.ctx = ops.context(self.device)
        self.sp = sp  # fairygen_b200.sp.SequenceParallel or None
        if sp is not None and cfg.num_heads % sp.world != 0:
            raise ValueError(f"num_heads {cfg.num_heads} not divisible by sequence-parallel world {sp.world}")
        self.rope_tab = torch.from_numpy(ops.rope_table(cfg.head_dim)).to(self.device)
        self._init_state()

    def _init_state(self) -> None:
        """Host-side state that does not touch the device (the CPU host-logic tests build an engine around it)."""
        self.blocks = []
        self._ws: Dict[Tuple[int, int], Dict[str, torch.Tensor]] = {}
        self._ctx_cache: Dict[tuple, tuple] = {}
        self._ctx_cache_order = []
        self.kernel_launches = 0  # kernels launched by forward() calls since the last reset
        self.ctx_cache_misses = 0        # contexts whose text embedding + 30 cross K|V projections were computed
        self.ctx_cache_content_hits = 0  # identity misses resolved by content (caller re-created an equal tensor)
        self.fused_adapters = []         # (lora_sd, alpha, targets) fused by lora_io.fuse_into_engine since the last pack
        self.timer = None         # optional fairygen_b200.profiling.KernelTimer
        self.loaded = False
        import os
        self.query_bounds = int(os.environ.get("FGB_QMAX", "0"))   # bit 0: self-attention, bit 1: cross-attention (see _workspace)

    # ------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Pack reference ``WanModel`` weights (state-dict names of DIT:271-336), already LoRA-fused
        if ``pipe.load_lora`` was applied (utils/lora/general.py:44-62), into kernel layouts."""
        cfg, dev = self.cfg, self.device
        g = lambda k: _dev_bf16(sd[k], dev)  # noqa: E731
        d = cfg.dim
        self.w_patch = g("patch_embedding.weight").reshape(d, -1).contiguous()  # [D, C*1*2*2], k = c*4 + y*2 + z
        self.b_patch = g("patch_embedding.bias")
        self.w_text0, self.b_text0 = g("text_embedding.0.weight"), g("text_embedding.0.bias")
        self.w_text2, self.b_text2 = g("text_embedding.2.weight"), g("text_embedding.2.bias")
        self.w_time0, self.b_time0 = g("time_embedding.0.weight"), g("time_embedding.0.bias")
        self.w_time2, self.b_time2 = g("time_embedding.2.weight"), g("time_embedding.2.bias")
        self.w_tproj, self.b_tproj = g("time_projection.1.weight"), g("time_projection.1.bias")
        self.w_head, self.b_head = g("head.head.weight"), g("head.head.bias")
        self.head_mod = g("head.modulation").reshape(1, 2 * d).contiguous()
        mods = []
        self.blocks = []
        for i in range(cfg.num_layers):
            p = f"blocks.{i}."
            b = _Block()
            b.private = {"wqkv", "cwkv"}
            sa, ca = p + "self_attn.", p + "cross_attn."
            b.wqkv = torch.cat([g(sa + "q.weight"), g(sa + "k.weight"), g(sa + "v.weight")], dim=0).contiguous()
            b.bqkv = torch.cat([g(sa + "q.bias"), g(sa + "k.bias"), g(sa + "v.bias")], dim=0).contiguous()
            b.wo, b.bo = g(sa + "o.weight"), g(sa + "o.bias")
            b.nq, b.nk = g(sa + "norm_q.weight"), g(sa + "norm_k.weight")
            b.cwq, b.cbq = g(ca + "q.weight"), g(ca + "q.bias")
            b.cwkv = torch.cat([g(ca + "k.weight"), g(ca + "v.weight")], dim=0).contiguous()
            b.cbkv = torch.cat([g(ca + "k.bias"), g(ca + "v.bias")], dim=0).contiguous()
            b.cwo, b.cbo = g(ca + "o.weight"), g(ca + "o.bias")
            b.cnq, b.cnk = g(ca + "norm_q.weight"), g(ca + "norm_k.weight")
            b.n3w, b.n3b = g(p + "norm3.weight"), g(p + "norm3.bias")
            b.w1, b.b1 = g(p + "ffn.0.weight"), g(p + "ffn.0.bias")
            b.w2, b.b2 = g(p + "ffn.2.weight"), g(p + "ffn.2.bias")
            mods.append(g(p + "modulation").reshape(1, 6 * d))
            self.blocks.append(b)
        self.mods_all = torch.cat(mods, dim=0).contiguous()  # [L, 6D]
        self._ctx_cache.clear()
        self._ctx_cache_order.clear()
        self.fused_adapters = []
        self.loaded = True

    _WEIGHT_ATTRS = ("w_patch", "b_patch", "w_text0", "b_text0", "w_text2", "b_text2", "w_time0", "b_time0", "w_time2", "b_time2",
                     "w_tproj", "b_tproj", "w_head", "b_head", "head_mod", "mods_all", "blocks")

    def share_weights_from(self, other: "WanDiTEngine") -> None:
        """Use `other`'s packed weight tensors (same device, same config) instead of packing a second 10 GB copy — e.g. one
        engine per parallel layout in one process.  Caches and workspaces stay private."""
        if not other.loaded or other.cfg != self.cfg or other.device != self.device:
            raise ValueError("share_weights_from needs a loaded engine of the same config on the same device")
        for name in self._WEIGHT_ATTRS:
            setattr(self, name, getattr(other, name))
        self._ctx_cache.clear()
        self._ctx_cache_order.clear()
        self.fused_adapters = other.fused_adapters
        self.loaded = True

    # ------------------------------------------------------------------------------------------
    # workspaces / caches
    # ------------------------------------------------------------------------------------------
    def _workspace(self, rows: int, s_pad: int) -> Dict[str, torch.Tensor]:
        key = (rows, s_pad)
        ws = self._ws.get(key)
        if ws is None:
            cfg, dev = self.cfg, self.device
            d, L = cfg.dim, cfg.num_layers
            e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)  # noqa: E731
            io = cfg.in_dim * 4
            ws = dict(
                x=e(rows, d), a=e(rows, d), qkv=e(rows, 3 * d), o=e(rows, d), cq=e(rows, d), h=e(rows, cfg.ffn_dim),
                prow=e(rows, io), hrow=e(rows, cfg.out_dim * 4),
                ts=torch.zeros(2, dtype=torch.float32, device=dev), emb=e(2, cfg.freq_dim), t0=e(2, d), t0s=e(2, d),
                t=e(2, d), ts_silu=e(2, d), tmod=e(2, 6 * d), mod_tab=e(2, L, 6 * d), head_tab=e(2, 2 * d),
                kmax2=torch.zeros(cfg.num_heads, dtype=torch.float32, device=dev),
                # head-level query bounds [self, cross] for fgb_attn_fwd_bounded_qk (FGB_QMAX bit 0 / bit 1; default off). Inside the
                # step the 16-lane sums that produce them cost the issue-bound norm kernels what the attention kernels gain:
                # cross-q rmsnorm +2.0-2.7 ms vs cross-attention -2.1 ms per step, qk_norm_rope +3.4 ms vs self-attention -1.7 ms
                # (profiles/r02_attn_head_bound_ab.log). A caller that has the bounds for free passes them to ops.attention.
                qmax2=torch.zeros(2, cfg.num_heads, dtype=torch.float32, device=dev),
                sk=ops.gemm_workspace(dev),    # stream-K tail of the block's GEMMs (all launches of a forward are on one stream)
            )
            if self.sp is not None:
                hpr_w = 3 * d // self.sp.world
                ws.update(hgather=e(s_pad, cfg.out_dim * 4))
                if getattr(self.sp, "exchange", "nccl") == "p2p":
                    ar = self.sp.peer_arena(rows, cfg.num_heads, dev)   # peers store straight into recv / o
                    ws.update(recv=ar.recv, o=ar.o)
                else:
                    ws.update(send=e(s_pad, hpr_w), recv=e(s_pad, hpr_w), o_full=e(s_pad, d // self.sp.world),
                              o_recv=e(s_pad, d // self.sp.world))
            self._ws = {key: ws}  # keep one shape resident (a video has one shape for all 100 forwards)
        return ws

    def _launched(self, n: int = 1) -> None:
        self.kernel_launches += n

    def _k(self, name, fn, *args, **kwargs):
        """Launch one kernel; under a KernelTimer bracket it with CUDA events on the launching stream."""
        self.kernel_launches += 1
        if self.timer is not None:
            return self.timer.call(name, fn, *args, **kwargs)
        return fn(*args, **kwargs)

    def _time_tables(self, ws, timestep: torch.Tensor, R: int):
        """Modulation tables for the R distinct timestep rows (row 0 = t=0 first-frame tokens when R == 2):
        mod_tab [R][L][6D] = block modulation + time_projection(t) (DIT:217-218), head_tab [R][2D] (DIT:263)."""
        dev, d = self.device, self.cfg.dim
        ts = ws["ts"]
        ts.zero_()
        ts[R - 1:R].copy_(timestep.reshape(-1)[:1].to(device=dev, dtype=torch.float32))
        emb, t0, t0s, t, tss, tmod = (ws[k][:R] for k in ("emb", "t0", "t0s", "t", "ts_silu", "tmod"))
        ops.sinusoidal_embedding(ts[:R], emb)
        ops.gemm(emb, self.w_time0, self.b_time0, t0)
        ops.silu(t0, t0s)
        ops.gemm(t0s, self.w_time2, self.b_time2, t)
        ops.silu(t, tss)
        ops.gemm(tss, self.w_tproj, self.b_tproj, tmod)
        mod_tab, head_tab = ws["mod_tab"], ws["head_tab"]
        for r in range(R):
            ops.add_bcast(self.mods_all, tmod[r], mod_tab[r])                 # modulation + t_mod   (DIT:217-218)
            ops.add_bcast(self.head_mod, t[r], head_tab[r:r + 1], period=d)   # head.modulation + t  (DIT:263)
        self._launched(6 + 2 * R)
        return mod_tab, head_tab

    def text_embedding(self, context: torch.Tensor) -> torch.Tensor:
        """text_embedding MLP on [n, text_dim] -> [n, dim] (DIT:307-311, PIPE:1236)."""
        cfg, dev = self.cfg, self.device
        c = context.reshape(-1, cfg.text_dim).to(device=dev, dtype=BF16).contiguous()
        hid = torch.empty(c.shape[0], cfg.dim, dtype=BF16, device=dev)
        emb = torch.empty(c.shape[0], cfg.dim, dtype=BF16, device=dev)
        ops.gemm(c, self.w_text0, self.b_text0, hid, EPI_BIAS_GELU_TANH)
        ops.gemm(hid, self.w_text2, self.b_text2, emb, EPI_BIAS)
        self._launched(2)
        return emb

    def _context_kv(self, context: torch.Tensor):
        """text_embedding + per-block cross-attention K (RMS-normed) | V; cached per context.

        Fast path: the caller passes the same tensor object every step (the reference pipeline does: ``inputs_posi["context"]``,
        PIPE:296) -> O(1) identity key (storage pointer, shape, dtype, version counter).  A caller that re-creates an equal
        tensor every step misses that key; the miss is then resolved by CONTENT against the cached contexts (one device
        compare of <= 4 MB + a sync, only on an identity miss) before 30 K|V GEMMs are re-done."""
        key = (context.data_ptr(), tuple(context.shape), context._version, context.dtype)
        hit = self._ctx_cache.get(key)
        if hit is not None:
            return hit
        for old_key in reversed(self._ctx_cache_order):
            cand = self._ctx_cache.get(old_key)
            if cand is None:
                continue
            ref = cand[2]
            if ref.shape == context.shape and ref.dtype == context.dtype and ref.device == context.device and torch.equal(ref, context):
                self.ctx_cache_content_hits += 1
                # cand[2] is the PRIVATE copy of the content (the caller may edit its tensor later); the 5th field keeps THIS
                # tensor alive so that its address cannot be recycled for other data while the identity key exists
                self._remember_context(key, cand[:4] + (context,))
                return cand
        cfg, dev = self.cfg, self.device
        d = cfg.dim
        self.ctx_cache_misses += 1
        emb = self.text_embedding(context)
        n = emb.shape[0]
        kv = torch.empty(cfg.num_layers, n, 2 * d, dtype=BF16, device=dev)
        kmax = torch.empty(cfg.num_layers, cfg.num_heads, dtype=torch.float32, device=dev)   # key bounds of the bounded softmax
        for i, b in enumerate(self.blocks):
            ops.gemm(emb, b.cwkv, b.cbkv, kv[i], EPI_BIAS)
            ops.rmsnorm_rope(kv[i][:, :d], cfg.eps, b.cnk)
            ops.head_norm_max(kv[i][:, :d], kmax[i], cfg.num_heads)
        self._launched(3 * cfg.num_layers)
        # keep a private copy of the context: it pins the content the entry was computed from (an in-place edit of the caller's
        # tensor bumps _version and misses the key) and cannot be freed / recycled under the key
        val = (kv, n, context.detach().clone(), kmax, context)
        self._remember_context(key, val)
        return val

    def _remember_context(self, key, val) -> None:
        self._ctx_cache[key] = val
        self._ctx_cache_order.append(key)
        while len(self._ctx_cache_order) > 8:
            self._ctx_cache.pop(self._ctx_cache_order.pop(0), None)

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, latents: torch.Tensor, timestep: torch.Tensor, context: torch.Tensor,
                fuse_vae_embedding_in_latents: bool = True, gather_output: bool = True) -> torch.Tensor:
        """latents (1,C,F,H,W); timestep (1,); context (1,L,text_dim) -> prediction shaped like latents."""
        if not self.loaded:
            raise RuntimeError("WanDiTEngine.forward called before load_state_dict")
        cfg = self.cfg
        if latents.dim() != 5 or latents.shape[0] != 1 or latents.shape[1] != cfg.in_dim:
            raise ValueError(f"latents must be (1,{cfg.in_dim},F,H,W), got {tuple(latents.shape)}")
        if latents.shape[3] % 2 or latents.shape[4] % 2:
            raise ValueError("latent height/width must be even (patch 2x2)")
        if not (context.dim() == 2 or (context.dim() == 3 and context.shape[0] == 1)) or context.shape[-1] != cfg.text_dim:
            raise ValueError(f"context must be (1,L,{cfg.text_dim}) or (L,{cfg.text_dim}), got {tuple(context.shape)}: a batch of "
                             "contexts (cfg_merge) is one forward per context — model_fn_wan_video does that split")
        dev, d, H = self.device, cfg.dim, cfg.num_heads
        f, h, w = latents.shape[2], latents.shape[3] // 2, latents.shape[4] // 2
        grid = (f, h, w)
        S = f * h * w
        sp = self.sp
        P, rank = (sp.world, sp.rank) if sp is not None else (1, 0)
        rows = -(-S // P)          # tokens per rank (the last rank is zero-padded, PIPE:1312-1315)
        s_pad = rows * P
        tok0 = rank * rows
        ws = self._workspace(rows, s_pad)
        lat = latents[0].to(device=dev, dtype=BF16).contiguous()

        # ---- time embedding on the distinct timestep rows only (PIPE:1218-1231, DIT:67-71, 312-318)
        per_token = bool(cfg.seperated_timestep and fuse_vae_embedding_in_latents)
        R = 2 if per_token else 1
        mod_tab, head_tab = self._time_tables(ws, timestep, R)
        r_main = R - 1                                   # row used by tokens after the first frame
        n_first = max(0, min(rows, h * w - tok0)) if per_token else 0   # tokens of this rank at t = 0

        kv_all, n_ctx, _, kmax_all = self._context_kv(context)[:4]

        # ---- patch embedding: im2row + GEMM (DIT:305, PIPE:1253-1261)
        x, a, qkv, o, cq, hbuf = ws["x"], ws["a"], ws["qkv"], ws["o"], ws["cq"], ws["h"]
        ops.patchify_rows(lat, ws["prow"], grid, tok0)
        ops.gemm(ws["prow"], self.w_patch, self.b_patch, x)
        self._launched(2)

        k = self._k
        sk = ws["sk"]
        qmax_self = ws["qmax2"][0] if self.query_bounds & 1 else None
        qmax_cross = ws["qmax2"][1] if self.query_bounds & 2 else None
        for i, b in enumerate(self.blocks):
            m0, m1 = mod_tab[0, i].view(6, d), mod_tab[r_main, i].view(6, d)
            # self-attention branch (DIT:224-225)
            k("ln_modulate", ops.ln_modulate, x, a, cfg.eps, m0[0], m0[1], m1[0], m1[1], n_first)
            if sp is not None and getattr(sp, "exchange", "nccl") == "p2p":
                def qkv_rows(r0, r1, a=a, b=b, qkv=qkv):
                    k("gemm_qkv", ops.gemm, a[r0:r1], b.wqkv, b.bqkv, qkv[r0:r1], sk_ws=sk)
                fused = (a, b.wqkv, b.bqkv, cfg.eps, b.nq, b.nk, self.rope_tab, grid) if (getattr(sp, "fused_send", False) and d % 256 == 0) else None
                sp.attention(self, ws, qkv, o, S, norm=(cfg.eps, b.nq, b.nk, self.rope_tab, grid, tok0), qkv_gemm=qkv_rows, fused=fused)
            else:
                k("gemm_qkv", ops.gemm, a, b.wqkv, b.bqkv, qkv, sk_ws=sk)
                if sp is None:
                    # q and k normalised + rotated and the key bound of the bounded softmax in one pass over the fused rows
                    k("rmsnorm_rope", ops.qk_norm_rope, qkv, d, cfg.eps, b.nq, b.nk, self.rope_tab, grid, tok0, ws["kmax2"], qmax_self)
                    k("attn_self", ops.attention, qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, H, kmax2=ws["kmax2"], qmax2=qmax_self)
                else:
                    k("rmsnorm_rope", ops.rmsnorm_rope, qkv[:, :d], cfg.eps, b.nq, self.rope_tab, grid, tok0)
                    k("rmsnorm_rope", ops.rmsnorm_rope, qkv[:, d:2 * d], cfg.eps, b.nk, self.rope_tab, grid, tok0)
                    sp.attention(self, ws, qkv, o, S)
            k("gemm_o", ops.gemm, o, b.wo, b.bo, x, EPI_GATED_RESIDUAL, m0[2], m1[2], n_first, sk_ws=sk)
            # cross-attention branch (DIT:226)
            k("ln_affine", ops.ln_affine, x, a, cfg.eps, b.n3w, b.n3b)
            k("gemm_cross_q", ops.gemm, a, b.cwq, b.cbq, cq, sk_ws=sk)
            k("rmsnorm", ops.rmsnorm_rope, cq, cfg.eps, b.cnq, hmax2=qmax_cross)
            k("attn_cross", ops.attention, cq, kv_all[i][:, :d], kv_all[i][:, d:], o, H, kmax2=kmax_all[i], qmax2=qmax_cross)
            k("gemm_cross_o", ops.gemm, o, b.cwo, b.cbo, x, EPI_RESIDUAL, sk_ws=sk)
            # feed-forward branch (DIT:227-228)
            k("ln_modulate", ops.ln_modulate, x, a, cfg.eps, m0[3], m0[4], m1[3], m1[4], n_first)
            k("gemm_ffn1", ops.gemm, a, b.w1, b.b1, hbuf, EPI_BIAS_GELU_TANH, sk_ws=sk)
            k("gemm_ffn2", ops.gemm, hbuf, b.w2, b.b2, x, EPI_GATED_RESIDUAL, m0[5], m1[5], n_first, sk_ws=sk)

        # ---- head (DIT:261-268) + unpatchify (DIT:346-351)
        h0, h1 = head_tab[0].view(2, d), head_tab[r_main].view(2, d)
        ops.ln_modulate(x, a, cfg.eps, h0[0], h0[1], h1[0], h1[1], n_first)
        ops.gemm(a, self.w_head, self.b_head, ws["hrow"])
        self._launched(2)
        hrow = ws["hrow"]
        if sp is not None:
            if not gather_output:
                return hrow
            hrow = sp.all_gather_rows(hrow, ws["hgather"])  # PIPE:1379-1382
        out = torch.empty(cfg.out_dim, f, 2 * h, 2 * w, dtype=BF16, device=dev)
        ops.unpatchify(hrow, out, grid)
        self._launched(1)
        return out.unsqueeze(0).to(latents.dtype)
