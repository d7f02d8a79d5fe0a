"""fairygen_b200 — B200-native (sm_100a) implementation of FairyGen's animation hot path: the
Wan2.2-TI2V-5B DiT denoising step behind the reference's ``pipe.model_fn`` surface.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + C ABI), the ctypes binding, the
engine that mirrors ``model_fn_wan_video``, the fused flow-match scheduler, the denoise loop, the
Ulysses / CFG / shot parallel layout (NVLink peer-store exchange), the stage-1/2 LoRA trainer
(hand-written backward), LoRA checkpoint formats, checkpoint detection / loading the umT5 text encoder and the VAE38 decoder / encoder.  There is no CPU
or PyTorch fallback.
"""
from .config import TI2V_5B, WanDiTConfig, counted_flops  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not require CUDA or the built library
    if name in ("WanDiTEngine",):
        from .engine import WanDiTEngine
        return WanDiTEngine
    if name in ("model_fn_wan_video", "install", "engine_for"):
        from . import model_fn
        return getattr(model_fn, name)
    if name == "FlowMatchScheduler":
        from .scheduler import FlowMatchScheduler
        return FlowMatchScheduler
    if name == "WanDenoiser":
        from .pipeline import WanDenoiser
        return WanDenoiser
    if name == "Stage2Trainer":
        from .training import Stage2Trainer
        return Stage2Trainer
    if name in ("UMT5Encoder", "UMT5Config", "UMT5_XXL"):
        from . import text_encoder
        return getattr(text_encoder, name)
    if name in ("VAE38Decoder", "VAE38Config", "VAE38"):
        from . import vae
        return getattr(vae, name)
    if name == "VAE38Encoder":
        from .vae_encode import VAE38Encoder
        return VAE38Encoder
    if name in ("lora_io", "cfg_parallel", "training", "text_encoder", "vae", "vae_encode"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name == "SequenceParallel":
        from .sp import SequenceParallel
        return SequenceParallel
    raise AttributeError(name)
