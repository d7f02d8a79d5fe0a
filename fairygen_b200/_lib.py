"""ctypes binding of the C ABI declared in ``include/fairygen_b200.h``.

The shared library is built in-tree (``fairygen_b200/libfairygen_b200.so``) by
``fairygen_b200/csrc/Makefile``.  There is no fallback: if the library is missing, or a call
returns non-zero, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FGB_LIB_PATH") or os.path.join(_HERE, "libfairygen_b200.so")   # override: A/B builds of one kernel
CSRC_DIR = os.path.join(_HERE, "csrc")

# name -> (restype, argtypes); must list every symbol of include/fairygen_b200.h
_P, _I32, _I64, _F = c_void_p, c_int32, c_int64, c_float
SIGNATURES = {
    "fgb_abi_version": (ctypes.c_int, []),
    "fgb_last_error": (c_char_p, []),
    "fgb_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_void_p)]),
    "fgb_destroy": (ctypes.c_int, [_P]),
    "fgb_sync_check": (ctypes.c_int, [_P, _P]),
    "fgb_sm_count": (ctypes.c_int, [_P]),
    "fgb_gemm_bf16": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _P]),
    "fgb_gemm_bf16_ex": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _P, _I64, _P, _I64, _I32, _P]),
    "fgb_gemm_dgrad_ex": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _P, _I64, _P, _I64, _I32, _P]),
    "fgb_attn_fwd": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _F, _P]),
    "fgb_attn_fwd_ex": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _F, _P, _I64, _P, _I64, _P]),
    "fgb_qk_norm_rope": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _F, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "fgb_rmsnorm_hmax": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _F, _P, _P, _P]),
    "fgb_attn_fwd_bounded_qk": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _F, _P, _P, _P, _I64, _P, _I64,
                                               ctypes.POINTER(c_void_p), _I32, _I32, _I32, _P]),
    "fgb_head_norm_max": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _P, _P]),
    "fgb_attn_set_stats": (ctypes.c_int, [_P, _P]),
    "fgb_attn_fwd_bounded": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _F, _P, _P, _I64, _P, _I64,
                                            ctypes.POINTER(c_void_p), _I32, _I32, _I32, _P]),
    "fgb_attn_bwd": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _I64, _P, _I64, _P, _I64, _P, _I64,
                                    _I32, _I32, _I32, _F, _P]),
    "fgb_attn_workspace_bytes": (c_int64, [_P, _I32, _I32, _I32]),
    "fgb_attn_schedule_check": (ctypes.c_int, [_I32, _I32, _I32, _I32, _I32, ctypes.POINTER(c_int32), ctypes.POINTER(c_int32)]),
    "fgb_ipc_export": (ctypes.c_int, [_P, _P, _P, ctypes.POINTER(c_int64)]),
    "fgb_ipc_open": (ctypes.c_int, [_P, _P, _I64, ctypes.POINTER(c_void_p)]),
    "fgb_ipc_close": (ctypes.c_int, [_P, _P, _I64]),
    "fgb_sp_scatter_heads": (ctypes.c_int, [_P, _P, _I64, ctypes.POINTER(c_void_p), _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "fgb_rmsnorm_rope_scatter": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _F, _P, _P, _I32, _I32, _I32, _I32, ctypes.POINTER(c_void_p),
                                                _I32, _I32, _I32, _I32, _P]),
    "fgb_sp_return_heads": (ctypes.c_int, [_P, _P, _I64, ctypes.POINTER(c_void_p), _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "fgb_sp_barrier": (ctypes.c_int, [_P, ctypes.POINTER(c_void_p), _I32, _I32, _I32, _P]),
    "fgb_sp_barrier_status": (ctypes.c_int, [_P, ctypes.POINTER(c_void_p), _I32, _I32, _I32, _P, _I64, _P]),
    "fgb_gemm_qkv_scatter": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I32, _I32, _I32, ctypes.POINTER(c_void_p), _I32, _I32, _P, _P, _I64,
                                            _P]),
    "fgb_gemm_bf16_sk": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _P, _I64, _P]),
    "fgb_gemm_workspace_bytes": (c_int64, [_P]),
    "fgb_gemm_streamk_tune": (ctypes.c_int, [_P, _I32, ctypes.c_double]),
    "fgb_gemm_schedule_check": (ctypes.c_int, [_I32, _I32, _I32, _I32, _I32, _I32, ctypes.POINTER(c_int32), ctypes.POINTER(c_int32)]),
    "fgb_sp_stats_barrier": (ctypes.c_int, [_P, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), _P, _I32, _I32, _P, _I32, _I32, _I32,
                                            _I32, _P, _P]),
    "fgb_recv_norm_rope": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _I32, _F, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P]),
    "fgb_attn_fwd_scatter": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, ctypes.POINTER(c_void_p), _I32, _I64, _I32, _I32, _I32, _I32,
                                            _I32, _F, _P, _I64, _P]),
    "fgb_gemm_dgrad": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _P]),
    "fgb_ln_bwd": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _F, _P, _P, _I32, _I32, _P]),
    "fgb_rmsnorm_rope_bwd": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _I32, _I32, _F, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "fgb_gelu_tanh": (ctypes.c_int, [_P, _P, _P, _I64, _P]),
    "fgb_gelu_tanh_bwd": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P]),
    "fgb_mul_gate": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _I32, _I32, _P, _P, _I32, _P]),
    "fgb_lora_merge": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _P, _P, _F, _F, _P, _I64, _I32, _I32, _I32, _P]),
    "fgb_lora_b2_eff_batched": (ctypes.c_int, [_P, _P, _P, _P, _I32, _I32, _F, _F, _P]),
    "fgb_lora_b2_eff": (ctypes.c_int, [_P, _P, _P, _F, _F, _P, _I64, _I64, _I32, _P]),
    "fgb_lora_wgrad": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _P, _F, _I32, _I32, _I32, _I32, _P]),
    "fgb_bernoulli_mask": (ctypes.c_int, [_P, _P, _I64, _F, ctypes.c_uint64, _P]),
    "fgb_fm_noise_target": (ctypes.c_int, [_P, _P, _P, _F, _P, _P, _I64, _P]),
    "fgb_mse_loss_grad": (ctypes.c_int, [_P, _P, _P, _F, _P, _P, _I64, _P]),
    "fgb_unpatchify_bwd": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P]),
    "fgb_adamw_step": (ctypes.c_int, [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I32, _P]),
    "fgb_conv_taps_bf16": (ctypes.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _I32,
                                          ctypes.POINTER(ctypes.c_int32), _I32, _I32, _I32, _P]),
    "fgb_vae_latent_rows": (ctypes.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "fgb_vae_norm_silu": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _I32, _P, _I32, _P]),
    "fgb_vae_upsample2x": (ctypes.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "fgb_vae_dup_up_add": (ctypes.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "fgb_vae_attn_softmax": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _F, _P]),
    "fgb_vae_unpatchify": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "fgb_vae_blend_finish": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "fgb_vae_patchify_rows": (ctypes.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "fgb_vae_space_to_depth": (ctypes.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "fgb_vae_avg_down_add": (ctypes.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "fgb_vae_latent_out": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _I32, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32,
                                          _I32, _I32, _P]),
    "fgb_vae_blend_divide": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "fgb_embedding_rows": (ctypes.c_int, [_P, _P, _I64, _I32, _P, _I32, _I32, _P, _I64, _P]),
    "fgb_t5_layer_norm": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _I32, _I32, _F, _P, _P]),
    "fgb_geglu": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _I32, _I32, _P]),
    "fgb_t5_bias_table": (ctypes.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P]),
    "fgb_t5_attention": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "fgb_ln_modulate": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _I32, _I32, _F, _P, _P, _P, _P, _I32, _P]),
    "fgb_ln_affine": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _I32, _I32, _F, _P, _P, _P]),
    "fgb_rmsnorm_rope": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _F, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "fgb_patchify_rows": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "fgb_unpatchify": (ctypes.c_int, [_P, _P, _I64, _P, _I32, _I32, _I32, _I32, _P]),
    "fgb_cfg_fm_step": (ctypes.c_int, [_P, _P, _P, _P, _P, _F, _F, _I32, _I32, _I32, _P]),
    "fgb_sinusoidal_embedding": (ctypes.c_int, [_P, _P, _P, _I32, _I32, _P]),
    "fgb_silu": (ctypes.c_int, [_P, _P, _P, _I64, _P]),
    "fgb_add_bcast": (ctypes.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _P]),
    "fgb_sp_pack_heads": (ctypes.c_int, [_P, _P, _I64, _P, _I32, _I32, _I32, _I32, _P]),
    "fgb_sp_unpack_heads": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P]),
}

EPI_BIAS, EPI_BIAS_GELU_TANH, EPI_GATED_RESIDUAL, EPI_RESIDUAL = 0, 1, 2, 3

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the shared library for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libfairygen_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the C-ABI library (once).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `make -C {CSRC_DIR}` or `python -c 'import __graft_entry__ as g; g.build()'`. "
                "fairygen_b200 has no CPU / PyTorch fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().fgb_last_error()
        raise RuntimeError(f"fairygen_b200 {what} failed (status {rc}): {msg.decode() if msg else '?'}")


class Context:
    """Owns one ``fgb_ctx`` (one per process/GPU)."""

    def __init__(self, device_index: int = 0):
        self._lib = lib()
        handle = c_void_p()
        check(self._lib.fgb_create(device_index, ctypes.byref(handle)), "fgb_create")
        self.handle = handle
        self.device_index = device_index
        self.sm_count = self._lib.fgb_sm_count(handle)

    def close(self):
        if getattr(self, "handle", None):
            self._lib.fgb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
