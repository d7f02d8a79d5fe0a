"""Torch-tensor front end of the C ABI (``include/fairygen_b200.h``).

PyTorch is used for device memory and streams only; every function here forwards raw device
pointers to a hand-written sm_100a kernel and raises on any non-zero status.  No function has a
PyTorch / CPU fallback.
"""
from __future__ import annotations

import math
from ctypes import c_void_p
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_BIAS_GELU_TANH, EPI_GATED_RESIDUAL, EPI_RESIDUAL  # noqa: F401

BF16 = torch.bfloat16
_ctx_by_device = {}


def context(device: torch.device) -> _lib.Context:
    """One fgb_ctx per CUDA device of this process."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"fairygen_b200 runs on CUDA (sm_100a) only, got device {device}; there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx_by_device:
        _ctx_by_device[idx] = _lib.Context(idx)
    return _ctx_by_device[idx]


_attn_stats = {}


def _stats_buffer(device) -> torch.Tensor:
    """Per-device int32[3] the attention kernel counts its CTAs into (fgb_attn_set_stats): [bound only, anchored, fallback]."""
    c = context(device)
    buf = _attn_stats.get(c.device_index)
    if buf is None:
        buf = torch.zeros(3, dtype=torch.int32, device=torch.device("cuda", c.device_index))
        _lib.check(_lib.lib().fgb_attn_set_stats(c.handle, c_void_p(buf.data_ptr())), "fgb_attn_set_stats")
        _attn_stats[c.device_index] = buf
    return buf


def attention_stats_reset(device) -> None:
    """Start counting the CTAs of bounded-score attention launches on `device` (and zero the counters)."""
    _stats_buffer(device).zero_()


def attention_stats(device):
    """(CTAs on the fixed-reference fast path [bound only + first-tile anchored], CTAs that fell back to the running-max path)
    since attention_stats_reset; (0, 0) if counting was never switched on."""
    c = context(device)
    buf = _attn_stats.get(c.device_index)
    if buf is None:
        return 0, 0
    a, b, f = (int(x) for x in buf.tolist())
    return a + b, f


def attention_stats_detail(device):
    buf = _stats_buffer(device)
    a, b, f = (int(x) for x in buf.tolist())
    return {"bound_only": a, "first_tile_anchored": b, "running_max_fallback": f}


def _h(t: torch.Tensor) -> _lib.Context:
    return context(t.device)


def _p(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _rowmajor(t: torch.Tensor, what: str) -> int:
    """Leading dimension (elements) of a 2-D bf16 view whose last dim is contiguous."""
    if t.dtype != BF16 or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{what}: expected a 2-D bf16 tensor with contiguous last dim, got {t.dtype} {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _vec(t: Optional[torch.Tensor], n: int, what: str) -> None:
    if t is not None and (t.dtype != BF16 or t.numel() != n or not t.is_contiguous()):
        raise ValueError(f"{what}: expected a contiguous bf16 vector of {n} elements")


def gemm_workspace(device) -> torch.Tensor:
    """Caller-owned, zero-initialised scratch of the stream-K tail (fgb_gemm_bf16_sk): one per engine; launches that share it must be
    ordered on one stream."""
    need = _lib.lib().fgb_gemm_workspace_bytes(context(device).handle)
    return torch.zeros(int(need), dtype=torch.uint8, device=device)


def gemm_streamk_tune(device, min_k: int = 6144, max_tail_frac: float = 0.9) -> None:
    """Thresholds of the stream-K split (fgb_gemm_streamk_tune); the defaults are the library's."""
    _lib.check(_lib.lib().fgb_gemm_streamk_tune(context(device).handle, min_k, max_tail_frac), "fgb_gemm_streamk_tune")


def gemm(a, w, bias, out, epilogue=EPI_BIAS, gate0=None, gate1=None, rows_gate0=0, a2=None, w2=None, sk_ws=None):
    """out[m,n] = epilogue(a[m,k] @ w[n,k].T (+ a2[m,k2] @ w2[n,k2].T) + bias); see fgb_gemm_bf16 / fgb_gemm_bf16_ex.
    sk_ws (gemm_workspace()): the last, partly filled wave of tiles is split along K (fgb_gemm_bf16_sk)."""
    lda, ldw, ldc = _rowmajor(a, "a"), _rowmajor(w, "w"), _rowmajor(out, "out")
    m, k = a.shape
    n = w.shape[0]
    if w.shape[1] != k or tuple(out.shape) != (m, n):
        raise ValueError(f"gemm shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)} out {tuple(out.shape)}")
    _vec(bias, n, "bias"), _vec(gate0, n, "gate0"), _vec(gate1, n, "gate1")
    c = _h(a)
    if sk_ws is not None and a2 is None and w2 is None:
        if sk_ws.dtype != torch.uint8 or not sk_ws.is_contiguous() or sk_ws.device != a.device:
            raise ValueError("gemm: sk_ws must be the uint8 tensor gemm_workspace() returned for this device")
        _lib.check(_lib.lib().fgb_gemm_bf16_sk(c.handle, _p(a), lda, _p(w), ldw, _p(bias), _p(out), ldc, m, n, k, epilogue,
                                               _p(gate0), _p(gate1), rows_gate0, _p(sk_ws), sk_ws.numel(), _stream()), "fgb_gemm_bf16_sk")
        return out
    if a2 is None and w2 is None:
        _lib.check(_lib.lib().fgb_gemm_bf16(c.handle, _p(a), lda, _p(w), ldw, _p(bias), _p(out), ldc, m, n, k, epilogue,
                                            _p(gate0), _p(gate1), rows_gate0, _stream()), "fgb_gemm_bf16")
        return out
    if a2 is None or w2 is None or a2.shape[0] != m or w2.shape[0] != n or a2.shape[1] != w2.shape[1]:
        raise ValueError("gemm: the second operand pair must be a2 [m, k2] and w2 [n, k2]")
    _lib.check(_lib.lib().fgb_gemm_bf16_ex(c.handle, _p(a), lda, _p(w), ldw, _p(bias), _p(out), ldc, m, n, k, epilogue, _p(gate0), _p(gate1),
                                           rows_gate0, _p(a2), _rowmajor(a2, "a2"), _p(w2), _rowmajor(w2, "w2"), a2.shape[1], _stream()),
               "fgb_gemm_bf16_ex")
    return out


_attn_ws = {}


def attention_workspace(q_rows: int, kv_rows: int, heads: int, device) -> Optional[torch.Tensor]:
    """Caller-owned scratch for the key-split tail of fgb_attn_fwd_ex (None when the grid needs no split)."""
    c = context(device)
    need = _lib.lib().fgb_attn_workspace_bytes(c.handle, q_rows, kv_rows, heads)
    if need <= 0:
        return None
    key = (c.device_index, torch.cuda.current_stream().cuda_stream)
    ws = _attn_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=torch.device("cuda", c.device_index))
        _attn_ws[key] = ws
    return ws


def _head_max(t, heads: int, name: str):
    if t is not None and (t.dtype != torch.float32 or t.numel() != heads or not t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous float32 [{heads}] tensor")


def qk_norm_rope(qkv, dim: int, eps: float, wq, wk, rope_tab, grid, token_offset: int, kmax2, qmax2=None):
    """In place on columns [0, 2*dim) of the fused q|k|v rows: RMSNorm + weight + RoPE of q and of k (one pass), and
    kmax2[h] = max over rows of ||k[row, h]||^2 (fp32 [dim/128]); qmax2 (optional) the same for q."""
    ld = _rowmajor(qkv, "qkv")
    if qkv.shape[1] < 2 * dim or kmax2.dtype != torch.float32 or kmax2.numel() != dim // 128 or not kmax2.is_contiguous():
        raise ValueError(f"qk_norm_rope: qkv {tuple(qkv.shape)} dim {dim} kmax2 {tuple(kmax2.shape)}")
    _head_max(qmax2, dim // 128, "qmax2")
    _vec(wq, dim, "wq")
    _vec(wk, dim, "wk")
    _lib.check(_lib.lib().fgb_qk_norm_rope(_h(qkv).handle, _p(qkv), ld, qkv.shape[0], dim, eps, _p(wq), _p(wk), _p(rope_tab), grid[0], grid[1],
                                           grid[2], token_offset, _p(kmax2), _p(qmax2), _stream()), "fgb_qk_norm_rope")


def head_norm_max(k, out_f32, heads: int):
    """out_f32[h] = max over rows of ||k[row, h*128:(h+1)*128]||^2 — the key bound of the bounded-score softmax."""
    ldk = _rowmajor(k, "k")
    if k.shape[1] != heads * 128 or out_f32.dtype != torch.float32 or out_f32.numel() != heads or not out_f32.is_contiguous():
        raise ValueError("head_norm_max: k [rows, heads*128] bf16, out float32 [heads]")
    _lib.check(_lib.lib().fgb_head_norm_max(_h(k).handle, _p(k), ldk, k.shape[0], heads, _p(out_f32), _stream()), "fgb_head_norm_max")
    return out_f32


def attention(q, k, v, out, heads: int, scale: Optional[float] = None, lse: Optional[torch.Tensor] = None,
              kmax2: Optional[torch.Tensor] = None, qmax2: Optional[torch.Tensor] = None):
    """out = softmax(q k^T scale) v per 128-wide head; q/out [s_q, heads*128], k/v [s_kv, heads*128].
    lse (optional, fp32 [heads, ld] with ld >= s_q, ld % 64 == 0) receives the log2-domain log-sum-exp rows
    for the backward pass."""
    ldq, ldk, ldv, ldo = _rowmajor(q, "q"), _rowmajor(k, "k"), _rowmajor(v, "v"), _rowmajor(out, "out")
    s_q, s_kv = q.shape[0], k.shape[0]
    width = heads * 128
    if q.shape[1] != width or k.shape[1] != width or v.shape != k.shape or out.shape != q.shape:
        raise ValueError(f"attention shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)} heads {heads}")
    scale = 1.0 / math.sqrt(128.0) if scale is None else scale
    if lse is not None and (lse.dtype != torch.float32 or lse.dim() != 2 or lse.shape[0] != heads or lse.shape[1] < s_q
                            or lse.shape[1] % 64 or not lse.is_contiguous()):
        raise ValueError("lse must be a contiguous float32 [heads, ld] tensor with ld >= s_q and ld % 64 == 0")
    c = _h(q)
    ws = attention_workspace(s_q, s_kv, heads, q.device)
    if kmax2 is not None:   # bounded-score softmax: fixed per-row reference from ||q_i||·max_j||k_j||, no running maximum
        if kmax2.dtype != torch.float32 or kmax2.numel() != heads or not kmax2.is_contiguous():
            raise ValueError("kmax2 must be a contiguous float32 [heads] tensor (head_norm_max of k)")
        _head_max(qmax2, heads, "qmax2")
        _lib.check(_lib.lib().fgb_attn_fwd_bounded_qk(c.handle, _p(q), ldq, _p(k), ldk, _p(v), ldv, _p(out), ldo, s_q, s_kv, heads, scale,
                                                      _p(kmax2), _p(qmax2), _p(lse), 0 if lse is None else lse.shape[1], _p(ws),
                                                      0 if ws is None else ws.numel(), None, 0, 0, 0, _stream()), "fgb_attn_fwd_bounded_qk")
        return out
    _lib.check(_lib.lib().fgb_attn_fwd_ex(c.handle, _p(q), ldq, _p(k), ldk, _p(v), ldv, _p(out), ldo, s_q, s_kv, heads,
                                          scale, _p(lse), 0 if lse is None else lse.shape[1], _p(ws),
                                          0 if ws is None else ws.numel(), _stream()), "fgb_attn_fwd_ex")
    return out


def stat_rows(s_q: int) -> int:
    """Row stride of the lse / delta buffers: s_q rounded up to the 64-row sub-tile of fgb_attn_bwd."""
    return -(-s_q // 64) * 64


def attention_bwd(q, k, v, out, dout, lse, dq, dk, dv, heads: int, scale: Optional[float] = None, delta=None):
    """Backward of :func:`attention`; writes dq, dk, dv (bf16, same layouts as q, k, v)."""
    s_q, s_kv = q.shape[0], k.shape[0]
    ld = [_rowmajor(t, n) for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out"), (dout, "dout"), (dq, "dq"), (dk, "dk"), (dv, "dv"))]
    width = heads * 128
    if any(t.shape[1] != width for t in (q, k, v, out, dout, dq, dk, dv)) or dq.shape != q.shape or dk.shape != k.shape or dv.shape != v.shape:
        raise ValueError("attention_bwd shape mismatch")
    if lse.dtype != torch.float32 or lse.dim() != 2 or lse.shape[0] != heads or lse.shape[1] % 64 or lse.shape[1] < s_q:
        raise ValueError("lse must be the float32 [heads, ld] tensor written by attention(..., lse=)")
    if delta is None:
        delta = torch.empty_like(lse)
    scale = 1.0 / math.sqrt(128.0) if scale is None else scale
    c = _h(q)
    _lib.check(_lib.lib().fgb_attn_bwd(c.handle, _p(q), ld[0], _p(k), ld[1], _p(v), ld[2], _p(out), ld[3], _p(dout), ld[4], _p(lse),
                                       _p(delta), lse.shape[1], _p(dq), ld[5], _p(dk), ld[6], _p(dv), ld[7], s_q, s_kv, heads,
                                       scale, _stream()), "fgb_attn_bwd")
    return dq, dk, dv


def ln_modulate(x, out, eps, shift0, scale0, shift1, scale1, rows_mod0):
    ldx, ldy = _rowmajor(x, "x"), _rowmajor(out, "out")
    rows, dim = x.shape
    for t in (shift0, scale0, shift1, scale1):
        _vec(t, dim, "modulation row")
    c = _h(x)
    _lib.check(_lib.lib().fgb_ln_modulate(c.handle, _p(x), ldx, _p(out), ldy, rows, dim, eps, _p(shift0), _p(scale0),
                                          _p(shift1), _p(scale1), rows_mod0, _stream()), "fgb_ln_modulate")
    return out


def ln_affine(x, out, eps, weight, bias):
    ldx, ldy = _rowmajor(x, "x"), _rowmajor(out, "out")
    rows, dim = x.shape
    _vec(weight, dim, "weight"), _vec(bias, dim, "bias")
    c = _h(x)
    _lib.check(_lib.lib().fgb_ln_affine(c.handle, _p(x), ldx, _p(out), ldy, rows, dim, eps, _p(weight), _p(bias),
                                        _stream()), "fgb_ln_affine")
    return out


def rmsnorm_rope(x, eps, weight, rope_tab=None, grid=(1, 1, 1), token_offset=0, hmax2=None):
    """In place on x [rows, dim] (may be a strided column slice of a wider buffer). hmax2 (fp32 [dim/128], no RoPE only):
    also leaves the per-head maximum of the squared norms of the output rows (fgb_rmsnorm_hmax)."""
    ldx = _rowmajor(x, "x")
    rows, dim = x.shape
    _vec(weight, dim, "weight")
    if hmax2 is not None:
        if rope_tab is not None:
            raise ValueError("rmsnorm_rope: hmax2 is for the un-rotated form (the cross-attention query)")
        _head_max(hmax2, dim // 128, "hmax2")
        _lib.check(_lib.lib().fgb_rmsnorm_hmax(_h(x).handle, _p(x), ldx, rows, dim, eps, _p(weight), _p(hmax2), _stream()), "fgb_rmsnorm_hmax")
        return x
    if rope_tab is not None and (rope_tab.dtype != torch.float32 or tuple(rope_tab.shape) != (1024, 64, 2) or not rope_tab.is_contiguous()):
        raise ValueError("rope_tab must be a contiguous float32 [1024, 64, 2] table")
    c = _h(x)
    _lib.check(_lib.lib().fgb_rmsnorm_rope(c.handle, _p(x), ldx, rows, dim, eps, _p(weight), _p(rope_tab), grid[0], grid[1],
                                           grid[2], token_offset, _stream()), "fgb_rmsnorm_rope")
    return x


def patchify_rows(latents, rows_out, grid, token_offset=0):
    """latents [C, f, 2h, 2w] bf16 contiguous -> rows_out [rows, >= 4C] (tokens token_offset..)."""
    if latents.dtype != BF16 or not latents.is_contiguous() or latents.dim() != 4:
        raise ValueError("latents must be a contiguous bf16 [C, F, H, W] tensor")
    ch = latents.shape[0]
    f, h, w = grid
    if tuple(latents.shape[1:]) != (f, 2 * h, 2 * w):
        raise ValueError(f"latents {tuple(latents.shape)} do not match grid {grid}")
    ld = _rowmajor(rows_out, "rows_out")
    c = _h(latents)
    _lib.check(_lib.lib().fgb_patchify_rows(c.handle, _p(latents), _p(rows_out), ld, ch, f, h, w, token_offset,
                                            rows_out.shape[0], _stream()), "fgb_patchify_rows")
    return rows_out


def unpatchify(head_rows, out, grid):
    """head_rows [f*h*w, 4C] -> out [C, f, 2h, 2w]."""
    ld = _rowmajor(head_rows, "head_rows")
    f, h, w = grid
    ch = out.shape[0]
    if out.dtype != BF16 or not out.is_contiguous() or tuple(out.shape) != (ch, f, 2 * h, 2 * w) or head_rows.shape[0] < f * h * w:
        raise ValueError("unpatchify shape mismatch")
    c = _h(out)
    _lib.check(_lib.lib().fgb_unpatchify(c.handle, _p(head_rows), ld, _p(out), ch, f, h, w, _stream()), "fgb_unpatchify")
    return out


def cfg_fm_step(latents, noise_pos, noise_neg, first_frame, cfg_scale: float, sigma_delta: float):
    """In place on latents [..., C, F, H, W] (batch 1): CFG + Euler step + first-frame restore."""
    lat = latents
    if lat.dtype != BF16 or not lat.is_contiguous():
        raise ValueError("latents must be contiguous bf16")
    ch, fr, hh, ww = lat.shape[-4:]
    for t in (noise_pos, noise_neg):
        if t is not None and (t.dtype != BF16 or not t.is_contiguous() or t.numel() != lat.numel()):
            raise ValueError("noise predictions must be contiguous bf16 tensors shaped like latents")
    if first_frame is not None and (first_frame.dtype != BF16 or not first_frame.is_contiguous() or first_frame.numel() != ch * hh * ww):
        raise ValueError("first_frame must be contiguous bf16 [C,1,H,W]")
    c = _h(lat)
    _lib.check(_lib.lib().fgb_cfg_fm_step(c.handle, _p(lat), _p(noise_pos), _p(noise_neg), _p(first_frame), cfg_scale,
                                          sigma_delta, ch, fr, hh * ww, _stream()), "fgb_cfg_fm_step")
    return latents


def sinusoidal_embedding(timesteps_f32, out):
    rows, dim = out.shape
    if timesteps_f32.dtype != torch.float32 or timesteps_f32.numel() != rows or out.dtype != BF16 or not out.is_contiguous():
        raise ValueError("sinusoidal_embedding: timesteps fp32 [rows], out bf16 [rows, dim]")
    c = _h(out)
    _lib.check(_lib.lib().fgb_sinusoidal_embedding(c.handle, _p(timesteps_f32), _p(out), rows, dim, _stream()),
               "fgb_sinusoidal_embedding")
    return out


def silu(x, out):
    if x.dtype != BF16 or out.dtype != BF16 or not x.is_contiguous() or not out.is_contiguous() or x.numel() != out.numel():
        raise ValueError("silu: contiguous bf16 tensors of equal size")
    c = _h(x)
    _lib.check(_lib.lib().fgb_silu(c.handle, _p(x), _p(out), x.numel(), _stream()), "fgb_silu")
    return out


def add_bcast(a, b, out, period: Optional[int] = None):
    """out[r, j] = a[r, j] + b[j % period]; a/out [rows, cols] contiguous bf16."""
    rows, cols = a.shape
    period = cols if period is None else period
    if a.dtype != BF16 or not a.is_contiguous() or out.shape != a.shape or not out.is_contiguous() or b.numel() < period or b.dtype != BF16:
        raise ValueError("add_bcast shape mismatch")
    c = _h(a)
    _lib.check(_lib.lib().fgb_add_bcast(c.handle, _p(a), _p(b), _p(out), rows, cols, period, _stream()), "fgb_add_bcast")
    return out


def sp_pack_heads(x, send, heads: int, groups: int, world: int):
    ldx = _rowmajor(x, "x")
    c = _h(x)
    _lib.check(_lib.lib().fgb_sp_pack_heads(c.handle, _p(x), ldx, _p(send), x.shape[0], heads, groups, world, _stream()),
               "fgb_sp_pack_heads")
    return send


def sp_unpack_heads(recv, x, heads: int, groups: int, world: int):
    ldx = _rowmajor(x, "x")
    c = _h(x)
    _lib.check(_lib.lib().fgb_sp_unpack_heads(c.handle, _p(recv), _p(x), ldx, x.shape[0], heads, groups, world, _stream()),
               "fgb_sp_unpack_heads")
    return x


def _ptr_array(ptrs):
    arr = (c_void_p * len(ptrs))(*[c_void_p(int(p)) for p in ptrs])
    return arr


def ipc_export(t: torch.Tensor):
    """(64-byte CUDA IPC handle, byte offset of t inside its allocation) for a caller-owned device tensor."""
    import ctypes
    handle = ctypes.create_string_buffer(64)
    off = ctypes.c_int64(0)
    _lib.check(_lib.lib().fgb_ipc_export(_h(t).handle, _p(t), handle, ctypes.byref(off)), "fgb_ipc_export")
    return bytes(handle.raw), int(off.value)


def ipc_open(device, handle: bytes, offset: int) -> int:
    import ctypes
    out = c_void_p()
    buf = ctypes.create_string_buffer(handle, 64)
    _lib.check(_lib.lib().fgb_ipc_open(context(device).handle, buf, offset, ctypes.byref(out)), "fgb_ipc_open")
    return int(out.value)


def ipc_close(device, peer_ptr: int, offset: int) -> None:
    _lib.check(_lib.lib().fgb_ipc_close(context(device).handle, c_void_p(peer_ptr), offset), "fgb_ipc_close")


def sp_scatter_heads(x, peer_ptrs, heads: int, groups: int, world: int, rank: int, group_first: int = 0, groups_total: Optional[int] = None):
    """Store this rank's [s_local, groups*heads*128] rows into groups [group_first, group_first+groups) of every peer's
    receive matrix (NVLink peer stores)."""
    ldx = _rowmajor(x, "x")
    c = _h(x)
    _lib.check(_lib.lib().fgb_sp_scatter_heads(c.handle, _p(x), ldx, _ptr_array(peer_ptrs), x.shape[0], heads, groups, group_first,
                                               groups if groups_total is None else groups_total, world, rank, _stream()),
               "fgb_sp_scatter_heads")


def rmsnorm_rope_scatter(x, eps, weight, rope_tab, grid, token_offset, peer_ptrs, world: int, rank: int, group: int, groups_total: int):
    """rmsnorm_rope(x) stored head by head into group `group` of the owning peers' receive matrices (x is not modified)."""
    ldx = _rowmajor(x, "x")
    rows, dim = x.shape
    _vec(weight, dim, "weight")
    c = _h(x)
    _lib.check(_lib.lib().fgb_rmsnorm_rope_scatter(c.handle, _p(x), ldx, rows, dim, eps, _p(weight), _p(rope_tab), grid[0], grid[1], grid[2],
                                                   token_offset, _ptr_array(peer_ptrs), world, rank, group, groups_total, _stream()),
               "fgb_rmsnorm_rope_scatter")


def sp_return_heads(x, peer_ptrs, ld_dst: int, rows: int, heads: int, groups: int, world: int, rank: int):
    """x [rows*world, groups*(heads/world)*128] (this rank's heads, all tokens) -> the token owners' [rows, groups*heads*128]."""
    ldx = _rowmajor(x, "x")
    _lib.check(_lib.lib().fgb_sp_return_heads(_h(x).handle, _p(x), ldx, _ptr_array(peer_ptrs), ld_dst, rows, x.shape[0], heads, groups,
                                              world, rank, _stream()), "fgb_sp_return_heads")


def gemm_qkv_scatter(a, w, bias, dim: int, peer_recv_ptrs, world: int, rank: int, rowsq, sk_ws=None):
    """q|k|v = a @ w.T + bias, every 32x64 block TMA-stored into the head owner's receive matrix (fused Ulysses send);
    rowsq fp32 [2, rows] += sum of squares of each q / k row."""
    lda, ldw = _rowmajor(a, "a"), _rowmajor(w, "w")
    m, k = a.shape
    if w.shape != (3 * dim, k) or rowsq.dtype != torch.float32 or rowsq.numel() != 2 * m or not rowsq.is_contiguous():
        raise ValueError(f"gemm_qkv_scatter: a {tuple(a.shape)} w {tuple(w.shape)} dim {dim} rowsq {tuple(rowsq.shape)}")
    _vec(bias, 3 * dim, "bias")
    _lib.check(_lib.lib().fgb_gemm_qkv_scatter(_h(a).handle, _p(a), lda, _p(w), ldw, _p(bias), m, dim, k, _ptr_array(peer_recv_ptrs), world,
                                               rank, _p(rowsq), _p(sk_ws), 0 if sk_ws is None else sk_ws.numel(), _stream()),
               "fgb_gemm_qkv_scatter")


def sp_stats_barrier(device, flag_ptrs, stats_ptrs, rowsq, rows: int, s_pad: int, kmax2, hpr: int, world: int, rank: int, epoch: int,
                     status=None):
    _lib.check(_lib.lib().fgb_sp_stats_barrier(context(device).handle, _ptr_array(flag_ptrs), _ptr_array(stats_ptrs), _p(rowsq), rows, s_pad,
                                               _p(kmax2), hpr, world, rank, epoch, _p(status), _stream()), "fgb_sp_stats_barrier")


def recv_norm_rope(recv, tokens: int, hpr: int, stats, dim: int, eps: float, wq, wk, rope_tab, grid, kmax2, qmax2=None):
    """RMSNorm (received full-row statistics) + weight slice + RoPE on the q, k groups of recv [s_pad, 3*hpr*128], in place;
    kmax2[h] = max ||k||^2 over the first `tokens` rows; qmax2 (optional) = max ||q||^2 over all rows."""
    if recv.dtype != BF16 or recv.dim() != 2 or recv.shape[1] != 3 * hpr * 128 or not recv.is_contiguous():
        raise ValueError(f"recv_norm_rope: recv {tuple(recv.shape)} hpr {hpr}")
    _lib.check(_lib.lib().fgb_recv_norm_rope(_h(recv).handle, _p(recv), recv.shape[0], tokens, hpr, _p(stats), dim, eps, _p(wq), _p(wk),
                                             _p(rope_tab), grid[0], grid[1], grid[2], _p(kmax2), _p(qmax2), _stream()), "fgb_recv_norm_rope")


def sp_barrier(device, flag_ptrs, world: int, rank: int, epoch: int, status: Optional[torch.Tensor] = None, timeout_clocks: int = 0):
    """Epoch barrier of the Ulysses exchange.  With ``status`` (device int32, zero while healthy) a peer that does not answer
    within ``timeout_clocks`` SM clocks (0 = ~30 s) is reported there and the kernel returns; without it the kernel traps."""
    if status is None:
        _lib.check(_lib.lib().fgb_sp_barrier(context(device).handle, _ptr_array(flag_ptrs), world, rank, epoch, _stream()), "fgb_sp_barrier")
        return
    if status.dtype != torch.int32 or status.numel() < 1:
        raise ValueError(f"sp_barrier: status must be a device int32, got {status.dtype} x {status.numel()}")
    _lib.check(_lib.lib().fgb_sp_barrier_status(context(device).handle, _ptr_array(flag_ptrs), world, rank, epoch, _p(status),
                                                int(timeout_clocks), _stream()), "fgb_sp_barrier_status")


def attention_scatter(q, k, v, o_peer_ptrs, ldo: int, rows_per_peer: int, col_offset: int, heads: int, scale: Optional[float] = None,
                      kmax2: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None, qmax2: Optional[torch.Tensor] = None):
    """attention() whose output rows go straight into the token-major o buffers of the ranks that own the tokens."""
    ldq, ldk, ldv = _rowmajor(q, "q"), _rowmajor(k, "k"), _rowmajor(v, "v")
    s_q, s_kv = q.shape[0], k.shape[0]
    if q.shape[1] != heads * 128 or k.shape[1] != heads * 128 or v.shape != k.shape:
        raise ValueError("attention_scatter shape mismatch")
    scale = 1.0 / math.sqrt(128.0) if scale is None else scale
    c = _h(q)
    ws = attention_workspace(s_q, s_kv, heads, q.device)
    if kmax2 is not None:
        _head_max(qmax2, heads, "qmax2")
        _lib.check(_lib.lib().fgb_attn_fwd_bounded_qk(c.handle, _p(q), ldq, _p(k), ldk, _p(v), ldv, None, ldo, s_q, s_kv, heads, scale,
                                                      _p(kmax2), _p(qmax2), _p(lse), 0 if lse is None else lse.shape[1], _p(ws),
                                                      0 if ws is None else ws.numel(), _ptr_array(o_peer_ptrs),
                                                      len(o_peer_ptrs), rows_per_peer, col_offset, _stream()), "fgb_attn_fwd_bounded_qk")
        return
    if lse is not None:
        raise ValueError("attention_scatter: lse needs the bounded form (pass kmax2)")
    _lib.check(_lib.lib().fgb_attn_fwd_scatter(c.handle, _p(q), ldq, _p(k), ldk, _p(v), ldv, _ptr_array(o_peer_ptrs), len(o_peer_ptrs), ldo,
                                               rows_per_peer, col_offset, s_q, s_kv, heads, scale, _p(ws), 0 if ws is None else ws.numel(),
                                               _stream()), "fgb_attn_fwd_scatter")


def rope_table(head_dim: int = 128, positions: int = 1024, theta: float = 10000.0) -> np.ndarray:
    """Host-side float32 [positions, head_dim/2, 2] (cos, sin) table for fgb_rmsnorm_rope.

    Same construction as the reference's precompute_freqs_cis_3d (DIT:74-88): the head's complex
    lanes are split frame/row/column = dim-2*(dim//3), dim//3, dim//3 real dims, each axis with its
    own theta^(-2i/axis_dim) ladder, evaluated in float64 and rounded once to float32.
    """
    axis_dims = (head_dim - 2 * (head_dim // 3), head_dim // 3, head_dim // 3)
    cols = []
    pos = np.arange(positions, dtype=np.float64)[:, None]
    for ad in axis_dims:
        inv = 1.0 / (theta ** (np.arange(0, ad, 2, dtype=np.float64)[: ad // 2] / ad))
        cols.append(pos * inv[None, :])
    ang = np.concatenate(cols, axis=1)  # [positions, head_dim/2]
    return np.stack([np.cos(ang), np.sin(ang)], axis=-1).astype(np.float32)


def sync_check(device=None):
    c = context(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    _lib.check(_lib.lib().fgb_sync_check(c.handle, _stream()), "fgb_sync_check")


# ---------------------------------------------------------------------------------------------------------
# training kernels (BASELINE config 5; see include/fairygen_b200.h "training")
# ---------------------------------------------------------------------------------------------------------
def gemm_dgrad(dy, w, dx, u=None, a1=None):
    """dx[m, in] = dy[m, out] @ w[out, in] (+ u[m, r] @ a1[r, in])  — input gradient of nn.Linear, W as stored."""
    ld_dy, ldw, ld_dx = _rowmajor(dy, "dy"), _rowmajor(w, "w"), _rowmajor(dx, "dx")
    m, k_out = dy.shape
    n_in = w.shape[1]
    if w.shape[0] != k_out or tuple(dx.shape) != (m, n_in):
        raise ValueError(f"gemm_dgrad shape mismatch: dy {tuple(dy.shape)} w {tuple(w.shape)} dx {tuple(dx.shape)}")
    c = _h(dy)
    if u is None and a1 is None:
        _lib.check(_lib.lib().fgb_gemm_dgrad(c.handle, _p(dy), ld_dy, _p(w), ldw, _p(dx), ld_dx, m, n_in, k_out, _stream()), "fgb_gemm_dgrad")
        return dx
    if u is None or a1 is None or u.shape[0] != m or a1.shape[1] != n_in or u.shape[1] != a1.shape[0]:
        raise ValueError("gemm_dgrad: the low-rank term must be u [m, r] and a1 [r, in]")
    _lib.check(_lib.lib().fgb_gemm_dgrad_ex(c.handle, _p(dy), ld_dy, _p(w), ldw, _p(dx), ld_dx, m, n_in, k_out, _p(u), _rowmajor(u, "u"),
                                            _p(a1), _rowmajor(a1, "a1"), u.shape[1], _stream()), "fgb_gemm_dgrad_ex")
    return dx


def ln_bwd(x, dy, out, eps, g0, g1=None, rows_mod0=0, affine=False, dres=None):
    rows, dim = x.shape
    _vec(g0, dim, "g0"), _vec(g1, dim, "g1")
    c = _h(x)
    _lib.check(_lib.lib().fgb_ln_bwd(c.handle, _p(x), _rowmajor(x, "x"), _p(dy), _rowmajor(dy, "dy"), _p(dres),
                                     0 if dres is None else _rowmajor(dres, "dres"), _p(out), _rowmajor(out, "out"), rows, dim, eps,
                                     _p(g0), _p(g1), rows_mod0, 1 if affine else 0, _stream()), "fgb_ln_bwd")
    return out


def rmsnorm_rope_bwd(x_pre, dy, eps, weight, rope_tab=None, grid=(1, 1, 1), token_offset=0):
    """In place on dy [rows, dim]: gradient w.r.t. the rotated, normalised output -> gradient w.r.t. x_pre."""
    rows, dim = x_pre.shape
    _vec(weight, dim, "weight")
    c = _h(dy)
    _lib.check(_lib.lib().fgb_rmsnorm_rope_bwd(c.handle, _p(x_pre), _rowmajor(x_pre, "x_pre"), _p(dy), _rowmajor(dy, "dy"), rows, dim, eps,
                                               _p(weight), _p(rope_tab), grid[0], grid[1], grid[2], token_offset, _stream()),
               "fgb_rmsnorm_rope_bwd")
    return dy


def _flat_bf16(*ts):
    n = ts[0].numel()
    for t in ts:
        if t.dtype != BF16 or not t.is_contiguous() or t.numel() != n:
            raise ValueError("expected contiguous bf16 tensors of equal size")
    return n


def gelu_tanh(z, h):
    n = _flat_bf16(z, h)
    _lib.check(_lib.lib().fgb_gelu_tanh(_h(z).handle, _p(z), _p(h), n, _stream()), "fgb_gelu_tanh")
    return h


def gelu_tanh_bwd(z, dh, dz):
    n = _flat_bf16(z, dh, dz)
    _lib.check(_lib.lib().fgb_gelu_tanh_bwd(_h(z).handle, _p(z), _p(dh), _p(dz), n, _stream()), "fgb_gelu_tanh_bwd")
    return dz


def mul_gate(dx, out, gate0, gate1, rows_gate0):
    rows, dim = dx.shape
    _vec(gate0, dim, "gate0"), _vec(gate1, dim, "gate1")
    _lib.check(_lib.lib().fgb_mul_gate(_h(dx).handle, _p(dx), _rowmajor(dx, "dx"), _p(out), _rowmajor(out, "out"), rows, dim,
                                       _p(gate0), _p(gate1), rows_gate0, _stream()), "fgb_mul_gate")
    return out


def lora_merge(w, a1, b1, b2, mask, w_eff, mask_mul=2.0, scaling=1.0):
    """w_eff = w + scaling * (b1 + bf16(bf16(b2*mask)*mask_mul)) @ a1;  w [n,k], a1 [r,k], b1/b2 [n,r], mask uint8 [n,r]."""
    n, k = w.shape
    r = a1.shape[0]
    for t, shp in ((a1, (r, k)), (b1, (n, r)), (b2, (n, r)), (w_eff, (n, k))):
        if t is not None and (tuple(t.shape) != shp or t.dtype != BF16):
            raise ValueError(f"lora_merge: expected bf16 {shp}, got {t.dtype} {tuple(t.shape)}")
    for t in (a1, b1, b2, mask):
        if t is not None and not t.is_contiguous():
            raise ValueError("lora_merge: a1, b1, b2 and mask must be contiguous")
    if mask is not None and (mask.dtype != torch.uint8 or tuple(mask.shape) != (n, r)):
        raise ValueError("lora_merge: mask must be uint8 [n, r]")
    _lib.check(_lib.lib().fgb_lora_merge(_h(w).handle, _p(w), _rowmajor(w, "w"), _p(a1), _rowmajor(a1, "a1"), _p(b1), _p(b2), _p(mask),
                                         mask_mul, scaling, _p(w_eff), _rowmajor(w_eff, "w_eff"), n, k, r, _stream()), "fgb_lora_merge")
    return w_eff


def lora_b2_eff_batched(b2_flat, mask_flat, table, rank: int, mask_mul=2.0, scaling=1.0):
    """lora_b2_eff for every Linear in one launch; table int64 [n, 4] on the device = (offset, rows, destination pointer, stride)."""
    if (b2_flat.dtype != BF16 or not b2_flat.is_contiguous() or table.dtype != torch.int64 or table.dim() != 2 or table.shape[1] != 4 or
            not table.is_contiguous() or table.device != b2_flat.device):
        raise ValueError("lora_b2_eff_batched: b2_flat contiguous bf16, table int64 [n, 4] on the same device")
    if mask_flat is not None and (mask_flat.dtype != torch.uint8 or mask_flat.numel() != b2_flat.numel() or not mask_flat.is_contiguous()):
        raise ValueError("lora_b2_eff_batched: mask_flat must be contiguous uint8 with one entry per B2 element")
    _lib.check(_lib.lib().fgb_lora_b2_eff_batched(_h(b2_flat).handle, _p(b2_flat), _p(mask_flat), _p(table), table.shape[0], rank, mask_mul,
                                                  scaling, _stream()), "fgb_lora_b2_eff_batched")


def lora_b2_eff(b2, mask, out, mask_mul=2.0, scaling=1.0):
    """out (view with contiguous last dim, e.g. a diagonal block) = scaling * bf16(bf16(b2 * mask) * mask_mul)."""
    n, r = b2.shape
    if b2.dtype != BF16 or not b2.is_contiguous() or tuple(out.shape) != (n, r) or out.dtype != BF16 or out.stride(1) != 1:
        raise ValueError("lora_b2_eff: b2 contiguous bf16 [n, r], out a bf16 [n, r] view")
    if mask is not None and (mask.dtype != torch.uint8 or tuple(mask.shape) != (n, r) or not mask.is_contiguous()):
        raise ValueError("lora_b2_eff: mask must be contiguous uint8 [n, r]")
    _lib.check(_lib.lib().fgb_lora_b2_eff(_h(b2).handle, _p(b2), _p(mask), mask_mul, scaling, _p(out), out.stride(0), n, r, _stream()),
               "fgb_lora_b2_eff")
    return out


def lora_wgrad(dy, t, db, mask=None, mul=1.0, transpose=False):
    """db[n, r] (fp32) += mul * mask * dy[s, n]^T @ t[s, r];  transpose=True: db is [r, n] (no mask)."""
    rows, n = dy.shape
    r = t.shape[1]
    want = (r, n) if transpose else (n, r)
    if t.shape[0] != rows or db.dtype != torch.float32 or tuple(db.shape) != want or not db.is_contiguous():
        raise ValueError("lora_wgrad shape mismatch")
    if mask is not None and (mask.dtype != torch.uint8 or tuple(mask.shape) != (n, r) or not mask.is_contiguous()):
        raise ValueError("lora_wgrad: mask must be contiguous uint8 [n, r]")
    _lib.check(_lib.lib().fgb_lora_wgrad(_h(dy).handle, _p(dy), _rowmajor(dy, "dy"), _p(t), _rowmajor(t, "t"), _p(db), _p(mask), mul,
                                         rows, n, r, 1 if transpose else 0, _stream()), "fgb_lora_wgrad")
    return db


def bernoulli_mask(out_u8, drop_prob: float, seed: int):
    if out_u8.dtype != torch.uint8 or not out_u8.is_contiguous():
        raise ValueError("bernoulli_mask: contiguous uint8 output")
    _lib.check(_lib.lib().fgb_bernoulli_mask(_h(out_u8).handle, _p(out_u8), out_u8.numel(), drop_prob, seed & (2 ** 64 - 1), _stream()),
               "fgb_bernoulli_mask")
    return out_u8


def fm_noise_target(x0, noise, sigma: float, latents, target):
    n = _flat_bf16(x0, noise, latents, target)
    _lib.check(_lib.lib().fgb_fm_noise_target(_h(x0).handle, _p(x0), _p(noise), sigma, _p(latents), _p(target), n, _stream()),
               "fgb_fm_noise_target")
    return latents, target


def mse_loss_grad(pred, target, weight: float, loss_f32, dpred=None):
    n = _flat_bf16(pred, target) if dpred is None else _flat_bf16(pred, target, dpred)
    if loss_f32.dtype != torch.float32 or loss_f32.numel() != 1:
        raise ValueError("mse_loss_grad: loss must be a float32 scalar tensor")
    _lib.check(_lib.lib().fgb_mse_loss_grad(_h(pred).handle, _p(pred), _p(target), weight, _p(loss_f32), _p(dpred), n, _stream()),
               "fgb_mse_loss_grad")
    return loss_f32


def unpatchify_bwd(dpred, d_rows, grid):
    f, h, w = grid
    ch = dpred.shape[0]
    if dpred.dtype != BF16 or not dpred.is_contiguous() or tuple(dpred.shape) != (ch, f, 2 * h, 2 * w) or d_rows.shape[0] < f * h * w:
        raise ValueError("unpatchify_bwd shape mismatch")
    _lib.check(_lib.lib().fgb_unpatchify_bwd(_h(dpred).handle, _p(dpred), _p(d_rows), _rowmajor(d_rows, "d_rows"), ch, f, h, w, _stream()),
               "fgb_unpatchify_bwd")
    return d_rows


def adamw_step(param, grad, m, v, lr, beta1, beta2, eps, weight_decay, step: int):
    n = param.numel()
    if param.dtype != BF16 or any(t.dtype != torch.float32 or t.numel() != n or not t.is_contiguous() for t in (grad, m, v)) or not param.is_contiguous():
        raise ValueError("adamw_step: bf16 param, fp32 grad/m/v of equal size")
    _lib.check(_lib.lib().fgb_adamw_step(_h(param).handle, _p(param), _p(grad), _p(m), _p(v), n, lr, beta1, beta2, eps, weight_decay, step,
                                         _stream()), "fgb_adamw_step")
    return param


# ---------------------------------------------------------------------------------------------------------------
# umT5 text encoder (fgb_embedding_rows, fgb_t5_layer_norm, fgb_geglu, fgb_t5_bias_table, fgb_t5_attention)
# ---------------------------------------------------------------------------------------------------------------
def embedding_rows(table, ids, out):
    """out[r] = table[ids[r]]; ids: int64 device vector."""
    if ids.dtype != torch.int64 or ids.dim() != 1 or not ids.is_contiguous() or ids.device != table.device:
        raise ValueError("embedding_rows: ids must be a contiguous int64 vector on the table's device")
    if out.shape[0] != ids.numel() or out.shape[1] != table.shape[1]:
        raise ValueError(f"embedding_rows shape mismatch: table {tuple(table.shape)} ids {tuple(ids.shape)} out {tuple(out.shape)}")
    _lib.check(_lib.lib().fgb_embedding_rows(_h(table).handle, _p(table), _rowmajor(table, "table"), table.shape[0], _p(ids), ids.numel(),
                                             table.shape[1], _p(out), _rowmajor(out, "out"), _stream()), "fgb_embedding_rows")
    return out


def t5_layer_norm(x, out, eps, weight):
    rows, dim = x.shape
    _vec(weight, dim, "weight")
    if weight is None or tuple(out.shape) != (rows, dim):
        raise ValueError("t5_layer_norm: weight is required and out must match x")
    _lib.check(_lib.lib().fgb_t5_layer_norm(_h(x).handle, _p(x), _rowmajor(x, "x"), _p(out), _rowmajor(out, "out"), rows, dim, eps,
                                            _p(weight), _stream()), "fgb_t5_layer_norm")
    return out


def geglu(gate_fc1, out):
    """out = gate_fc1[:, F:] * gelu_tanh(gate_fc1[:, :F])."""
    rows, two_f = gate_fc1.shape
    if two_f % 2 or tuple(out.shape) != (rows, two_f // 2):
        raise ValueError(f"geglu shape mismatch: gate_fc1 {tuple(gate_fc1.shape)} out {tuple(out.shape)}")
    _lib.check(_lib.lib().fgb_geglu(_h(out).handle, _p(gate_fc1), _rowmajor(gate_fc1, "gate_fc1"), _p(out), _rowmajor(out, "out"), rows,
                                    two_f // 2, _stream()), "fgb_geglu")
    return out


def t5_bias_table(emb, bucket_of_rel, out):
    """out [heads, n_rel] fp32 = emb[bucket_of_rel].T; emb [buckets, heads] bf16, bucket_of_rel int32 [n_rel]."""
    buckets, heads = emb.shape
    n_rel = bucket_of_rel.numel()
    if (emb.dtype != BF16 or not emb.is_contiguous() or bucket_of_rel.dtype != torch.int32 or not bucket_of_rel.is_contiguous() or
            out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (heads, n_rel)):
        raise ValueError("t5_bias_table: emb bf16 [buckets, heads], bucket_of_rel int32 [n_rel], out fp32 [heads, n_rel]")
    _lib.check(_lib.lib().fgb_t5_bias_table(_h(emb).handle, _p(emb), _p(bucket_of_rel), n_rel, heads, buckets, _p(out), _stream()),
               "fgb_t5_bias_table")
    return out


def t5_attention(q, k, v, out, batch: int, heads: int, bias=None, key_mask=None):
    """T5 attention core on [batch*s, heads*64] matrices (no score scaling); bias fp32 [heads, s_q+s_kv-1], key_mask uint8 [batch, s_kv]."""
    if q.shape[0] % batch or k.shape[0] % batch or k.shape[0] != v.shape[0] or out.shape[0] != q.shape[0]:
        raise ValueError("t5_attention: row counts must be batch * sequence length")
    s_q, s_kv = q.shape[0] // batch, k.shape[0] // batch
    if any(t.shape[1] != heads * 64 for t in (q, k, v, out)):
        raise ValueError("t5_attention: head_dim is 64 — every matrix must be heads*64 wide")
    if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous() or tuple(bias.shape) != (heads, s_q + s_kv - 1)):
        raise ValueError(f"t5_attention: bias must be fp32 [heads, s_q+s_kv-1] = [{heads}, {s_q + s_kv - 1}]")
    if key_mask is not None and (key_mask.dtype != torch.uint8 or not key_mask.is_contiguous() or tuple(key_mask.shape) != (batch, s_kv)):
        raise ValueError(f"t5_attention: key_mask must be uint8 [batch, s_kv] = [{batch}, {s_kv}]")
    _lib.check(_lib.lib().fgb_t5_attention(_h(q).handle, _p(q), _rowmajor(q, "q"), _p(k), _rowmajor(k, "k"), _p(v), _rowmajor(v, "v"),
                                           _p(out), _rowmajor(out, "out"), batch, s_q, s_kv, heads, _p(bias), _p(key_mask), _stream()),
               "fgb_t5_attention")
    return out


# ---------------------------------------------------------------------------------------------------------------
# VAE38 decoder (fgb_conv_taps_bf16, fgb_vae_*): feature maps are 2-D row views [T*(H+2)*(W+2), Cp] of zero-bordered grids
# ---------------------------------------------------------------------------------------------------------------
def conv_taps(x, a_row0: int, w, bias, out, tap_offsets, grid_hw=(0, 0), epilogue=EPI_BIAS):
    """out[r] = epilogue(sum_tap x[a_row0 + r + tap_offsets[tap]] @ w[:, tap*cin:(tap+1)*cin].T + bias) for r < out.shape[0]."""
    import ctypes

    ldx, ldw, ldo = _rowmajor(x, "x"), _rowmajor(w, "w"), _rowmajor(out, "out")
    cin, taps = x.shape[1], len(tap_offsets)
    m, n = out.shape
    if w.shape[0] != n or w.shape[1] != taps * cin:
        raise ValueError(f"conv_taps shape mismatch: x {tuple(x.shape)} w {tuple(w.shape)} out {tuple(out.shape)} taps {taps}")
    _vec(bias, n, "bias")
    offs = (ctypes.c_int32 * taps)(*[int(v) for v in tap_offsets])
    _lib.check(_lib.lib().fgb_conv_taps_bf16(_h(x).handle, _p(x), ldx, x.shape[0], a_row0, _p(w), ldw, _p(bias), _p(out), ldo, m, n, cin, taps,
                                             offs, grid_hw[0], grid_hw[1], epilogue, _stream()), "fgb_conv_taps_bf16")
    return out


def vae_latent_rows(z, mean, inv_std, grid, cp: int):
    """z bf16 [C, T, H, W] -> interior of grid (rows view of [T, H+2, W+2, cp]) = z / inv_std + mean."""
    C, T, H, W = z.shape
    if z.dtype != BF16 or not z.is_contiguous() or grid.shape[0] != T * (H + 2) * (W + 2) or grid.shape[1] != cp or not grid.is_contiguous():
        raise ValueError("vae_latent_rows: z contiguous bf16 [C, T, H, W], grid contiguous rows [T*(H+2)*(W+2), cp]")
    if any(t.dtype != torch.float32 or t.numel() != C or not t.is_contiguous() for t in (mean, inv_std)):
        raise ValueError("vae_latent_rows: mean / inv_std must be contiguous fp32 [C]")
    _lib.check(_lib.lib().fgb_vae_latent_rows(_h(z).handle, _p(z), _p(mean), _p(inv_std), _p(grid), C, T, H, W, cp, _stream()),
               "fgb_vae_latent_rows")


def vae_norm_silu(x, out, channels: int, gamma, silu: bool = True):
    rows, cp = x.shape
    if not x.is_contiguous() or not out.is_contiguous() or tuple(out.shape) != (rows, cp) or x.dtype != BF16 or out.dtype != BF16:
        raise ValueError("vae_norm_silu: x and out must be contiguous bf16 row matrices of one shape")
    _vec(gamma, cp, "gamma")
    _lib.check(_lib.lib().fgb_vae_norm_silu(_h(x).handle, _p(x), _p(out), rows, channels, cp, _p(gamma), 1 if silu else 0, _stream()),
               "fgb_vae_norm_silu")


def vae_upsample2x(src, dst, cp: int, frames_dst: int, h: int, w: int, halves: int = 1):
    if src.dtype != BF16 or dst.dtype != BF16 or not src.is_contiguous() or not dst.is_contiguous():
        raise ValueError("vae_upsample2x: contiguous bf16 grids")
    if src.numel() < (frames_dst // halves) * (h + 2) * (w + 2) * halves * cp or dst.numel() < frames_dst * (2 * h + 2) * (2 * w + 2) * cp:
        raise ValueError("vae_upsample2x: grid too small")
    _lib.check(_lib.lib().fgb_vae_upsample2x(_h(src).handle, _p(src), _p(dst), cp, frames_dst, h, w, halves, _stream()), "fgb_vae_upsample2x")


def vae_dup_up_add(x, main, cin: int, cout: int, factor_t: int, first_chunk: bool, frames_out: int, h: int, w: int):
    if x.dtype != BF16 or main.dtype != BF16 or not x.is_contiguous() or not main.is_contiguous():
        raise ValueError("vae_dup_up_add: contiguous bf16 grids")
    frames_in = (frames_out + (factor_t - 1 if first_chunk else 0)) // factor_t
    if x.shape[0] < frames_in * (h + 2) * (w + 2) or main.shape[0] < frames_out * (2 * h + 2) * (2 * w + 2):
        raise ValueError("vae_dup_up_add: grid too small")
    _lib.check(_lib.lib().fgb_vae_dup_up_add(_h(x).handle, _p(x), _p(main), cin, x.shape[1], cout, main.shape[1], factor_t,
                                             1 if first_chunk else 0, frames_out, h, w, _stream()), "fgb_vae_dup_up_add")


def vae_attn_softmax(scores, n_cols: int, grid_h: int, grid_w: int, scale: float):
    ld = _rowmajor(scores, "scores")
    _lib.check(_lib.lib().fgb_vae_attn_softmax(_h(scores).handle, _p(scores), ld, scores.shape[0], n_cols, grid_h, grid_w, scale, _stream()),
               "fgb_vae_attn_softmax")


def vae_unpatchify(head, frames: int, h: int, w: int, values, weight, t0: int, y0: int, x0: int, bounds=(True, True, True, True),
                   border=(1, 1)):
    """head rows [frames*(h+2)*(w+2), cp] -> values fp32 [3, VT, VH, VW] (+ weight fp32 [VT, VH, VW] when blending tiles)."""
    if head.dtype != BF16 or not head.is_contiguous() or head.shape[0] < frames * (h + 2) * (w + 2):
        raise ValueError("vae_unpatchify: head must be the contiguous bf16 rows of the head grid")
    if values.dtype != torch.float32 or values.dim() != 4 or values.shape[0] != 3 or not values.is_contiguous():
        raise ValueError("vae_unpatchify: values must be contiguous fp32 [3, T, H, W]")
    if weight is not None and (weight.dtype != torch.float32 or tuple(weight.shape) != tuple(values.shape[1:]) or not weight.is_contiguous()):
        raise ValueError("vae_unpatchify: weight must be contiguous fp32 [T, H, W]")
    b = (int(bounds[0]) << 3) | (int(bounds[1]) << 2) | (int(bounds[2]) << 1) | int(bounds[3])
    _lib.check(_lib.lib().fgb_vae_unpatchify(_h(head).handle, _p(head), frames, h, w, head.shape[1], _p(values), _p(weight), t0, y0, x0,
                                             values.shape[1], values.shape[2], values.shape[3], b, border[0], border[1], _stream()),
               "fgb_vae_unpatchify")


def vae_blend_finish(values, weight):
    _lib.check(_lib.lib().fgb_vae_blend_finish(_h(values).handle, _p(values), _p(weight), weight.numel(), values.shape[0], _stream()),
               "fgb_vae_blend_finish")


# ---- VAE38 encoder side ----
def vae_patchify_rows(video, grid, cp: int):
    """video bf16 [3, T, H, W] -> interior of grid rows [T*(H/2+2)*(W/2+2), cp]."""
    C, T, H, W = video.shape
    if video.dtype != BF16 or not video.is_contiguous() or C != 3 or H % 2 or W % 2:
        raise ValueError("vae_patchify_rows: video must be contiguous bf16 [3, T, H, W] with even H, W")
    if grid.dtype != BF16 or not grid.is_contiguous() or tuple(grid.shape) != (T * (H // 2 + 2) * (W // 2 + 2), cp):
        raise ValueError("vae_patchify_rows: grid must be contiguous bf16 rows [T*(H/2+2)*(W/2+2), cp]")
    _lib.check(_lib.lib().fgb_vae_patchify_rows(_h(video).handle, _p(video), _p(grid), T, H, W, cp, _stream()), "fgb_vae_patchify_rows")


def vae_space_to_depth(src, dst, cp: int, frames: int, h: int, w: int):
    if src.dtype != BF16 or dst.dtype != BF16 or not src.is_contiguous() or not dst.is_contiguous() or h % 2 or w % 2:
        raise ValueError("vae_space_to_depth: contiguous bf16 grids, even h and w")
    if src.numel() < frames * (h + 2) * (w + 2) * cp or dst.numel() < frames * (h // 2 + 2) * (w // 2 + 2) * 4 * cp:
        raise ValueError("vae_space_to_depth: grid too small")
    _lib.check(_lib.lib().fgb_vae_space_to_depth(_h(src).handle, _p(src), _p(dst), cp, frames, h, w, _stream()), "fgb_vae_space_to_depth")


def vae_avg_down_add(x, main, cin: int, cout: int, factor_t: int, factor_s: int, pad_front: int, frames_out: int, h_out: int, w_out: int):
    if x.dtype != BF16 or main.dtype != BF16 or not x.is_contiguous() or not main.is_contiguous():
        raise ValueError("vae_avg_down_add: contiguous bf16 grids")
    frames_in = frames_out * factor_t - pad_front
    if x.shape[0] < frames_in * (h_out * factor_s + 2) * (w_out * factor_s + 2) or main.shape[0] < frames_out * (h_out + 2) * (w_out + 2):
        raise ValueError("vae_avg_down_add: grid too small")
    _lib.check(_lib.lib().fgb_vae_avg_down_add(_h(x).handle, _p(x), _p(main), cin, x.shape[1], cout, main.shape[1], factor_t, factor_s,
                                               pad_front, frames_out, h_out, w_out, _stream()), "fgb_vae_avg_down_add")


def vae_latent_out(grid, frames: int, h: int, w: int, mean, inv_std, values, weight, t0: int, y0: int, x0: int,
                   bounds=(True, True, True, True), border=(1, 1)):
    """conv1 grid rows [frames*(h+2)*(w+2), cp] -> normalised latent mean into values fp32 [z, T, H, W] (+ weight when blending)."""
    z = values.shape[0]
    if grid.dtype != BF16 or not grid.is_contiguous() or grid.shape[0] < frames * (h + 2) * (w + 2) or grid.shape[1] < z:
        raise ValueError("vae_latent_out: grid must be the contiguous bf16 rows of the conv1 grid")
    if values.dtype != torch.float32 or values.dim() != 4 or not values.is_contiguous():
        raise ValueError("vae_latent_out: values must be contiguous fp32 [z, T, H, W]")
    if weight is not None and (weight.dtype != torch.float32 or tuple(weight.shape) != tuple(values.shape[1:]) or not weight.is_contiguous()):
        raise ValueError("vae_latent_out: weight must be contiguous fp32 [T, H, W]")
    if any(t.dtype != torch.float32 or t.numel() != z or not t.is_contiguous() for t in (mean, inv_std)):
        raise ValueError("vae_latent_out: mean / inv_std must be contiguous fp32 [z]")
    b = (int(bounds[0]) << 3) | (int(bounds[1]) << 2) | (int(bounds[2]) << 1) | int(bounds[3])
    _lib.check(_lib.lib().fgb_vae_latent_out(_h(grid).handle, _p(grid), frames, h, w, grid.shape[1], _p(mean), _p(inv_std), z, _p(values),
                                             _p(weight), t0, y0, x0, values.shape[1], values.shape[2], values.shape[3], b, border[0], border[1],
                                             _stream()), "fgb_vae_latent_out")


def vae_blend_divide(values, weight):
    _lib.check(_lib.lib().fgb_vae_blend_divide(_h(values).handle, _p(values), _p(weight), weight.numel(), values.shape[0], _stream()),
               "fgb_vae_blend_divide")
