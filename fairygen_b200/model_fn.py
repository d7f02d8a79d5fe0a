"""Drop-in for ``pipe.model_fn`` = ``model_fn_wan_video`` (reference animation/diffsynth/pipelines/
wan_video.py:1122-1388), TI2V-5B branches, backed by the sm_100a kernels.

Boundary: the reference calls ``self.model_fn(**models, **inputs_shared,
**inputs_posi, timestep=timestep)`` (wan_video.py:296, 301; diffusion/loss.py:17).  This function
honours that keyword signature, swallows the ~40 unrelated keys in ``**kwargs`` and returns the
velocity prediction with the shape/dtype/device of ``latents``.  The reference ``WanModel`` stays
the weight container (so ``pipe.load_lora(pipe.dit, ...)`` keeps working); the engine re-packs from
it whenever a parameter's version counter changes.

    import fairygen_b200
    fairygen_b200.install(pipe)        # pipe.model_fn -> this file; pipe.scheduler -> fused step
    video = pipe(prompt=..., input_image=..., ...)   # unchanged reference call

Anything outside the TI2V-5B inference path (VACE, S2V audio, Animate, VAP, LongCat, camera/motion
controllers, reference latents, TeaCache, sliding windows, CLIP/VAE-concat image inputs) raises
``NotImplementedError`` instead of silently falling back; so does a call under autograd with trainable
parameters — LoRA training goes through ``fairygen_b200.training.Stage2Trainer`` (same kernels + hand-written backward).
"""
from __future__ import annotations

from typing import Optional

import torch

from .config import WanDiTConfig
from .engine import WanDiTEngine

_ENGINE_ATTR = "_fairygen_b200_engine"
_VERSION_ATTR = "_fairygen_b200_weight_version"
_PARAMS_ATTR = "_fairygen_b200_param_list"


def _tracked_params(dit):
    """The module's parameters as a cached tuple.  ``load_state_dict(assign=True)`` / ``to_empty`` replace Parameter
    objects, so a load-state-dict post hook on the container and on every parameter-owning submodule (the reference's LoRA
    loader calls ``load_state_dict`` on the adapted Linear itself, LORA:60) drops the cache."""
    cached = getattr(dit, _PARAMS_ATTR, None)
    if cached is None:
        cached = tuple(dit.parameters())
        object.__setattr__(dit, _PARAMS_ATTR, cached)
        if not getattr(dit, "_fairygen_b200_hooked", False):
            def _invalidate(module, incompatible_keys):
                object.__setattr__(dit, _PARAMS_ATTR, None)

            for m in dit.modules():
                if m is dit or any(True for _ in m.parameters(recurse=False)):
                    m.register_load_state_dict_post_hook(_invalidate)
            object.__setattr__(dit, "_fairygen_b200_hooked", True)
    return cached


def _weights_version(dit):
    """Fingerprint of the container's weights: in-place updates bump ``_version`` (``load_state_dict``, ``pipe.load_lora``,
    optimizer steps), ``param.data = ...`` swaps and re-allocations change ``data_ptr``.  One pass over a cached tuple
    (~0.1 ms for the 825 tensors of TI2V-5B, against a 440 ms forward); not detected: writes through a detached alias of
    the storage (``p.data.add_()``) — call ``engine.load_state_dict(dit.state_dict())`` after those."""
    ver = 0
    ptr = 0
    for p in _tracked_params(dit):
        ver += p._version
        ptr ^= p.data_ptr()
    return ver, ptr, len(getattr(dit, _PARAMS_ATTR))


def _require_device_weights(device) -> None:
    if device.type != "cuda":
        raise RuntimeError(f"fairygen_b200: the DiT's weights are on `{device}`; the B200 path packs them from device memory — "
                           "move the model to the GPU first (the reference's offload / vram-management modes are outside the "
                           "hot path: disable them or call pipe.load_models_to_device(['dit']) before the first step)")


def engine_for(dit, sp=None) -> WanDiTEngine:
    """Engine attached to a reference ``WanModel``; packs (or re-packs after ``load_lora``) lazily.  Adapters fused directly
    into the packed weights (``lora_io.fuse_into_engine``) are re-applied after a re-pack."""
    eng: Optional[WanDiTEngine] = getattr(dit, _ENGINE_ATTR, None)
    version = _weights_version(dit)
    if eng is None or (sp is not None and eng.sp is not sp):
        device = next(dit.parameters()).device
        _require_device_weights(device)
        eng = WanDiTEngine(WanDiTConfig.from_module(dit), device=device, sp=sp)
        object.__setattr__(dit, _ENGINE_ATTR, eng)
        object.__setattr__(dit, _VERSION_ATTR, None)
    if getattr(dit, _VERSION_ATTR, None) != version:
        adapters = list(eng.fused_adapters)
        eng.load_state_dict(dit.state_dict())
        if adapters:
            from .lora_io import fuse_into_engine

            for lora_sd, alpha, targets in adapters:
                fuse_into_engine(eng, lora_sd, alpha, targets, _record=False)
            eng.fused_adapters = adapters
        object.__setattr__(dit, _VERSION_ATTR, version)
    return eng


def _reject(name: str, value) -> None:
    if value is not None:
        raise NotImplementedError(f"fairygen_b200 model_fn: `{name}` is outside the Wan2.2-TI2V-5B hot path (SURVEY.md §2, out of scope)")


def model_fn_wan_video(
    dit=None,
    motion_controller=None,
    vace=None,
    vap=None,
    animate_adapter=None,
    latents: torch.Tensor = None,
    timestep: torch.Tensor = None,
    context: torch.Tensor = None,
    clip_feature=None,
    y=None,
    reference_latents=None,
    vace_context=None,
    vace_scale=1.0,
    audio_embeds=None,
    motion_latents=None,
    s2v_pose_latents=None,
    vap_hidden_state=None,
    vap_clip_feature=None,
    context_vap=None,
    drop_motion_frames: bool = True,
    tea_cache=None,
    use_unified_sequence_parallel: bool = False,
    motion_bucket_id=None,
    pose_latents=None,
    face_pixel_values=None,
    longcat_latents=None,
    sliding_window_size=None,
    sliding_window_stride=None,
    cfg_merge: bool = False,
    use_gradient_checkpointing: bool = False,
    use_gradient_checkpointing_offload: bool = False,
    control_camera_latents_input=None,
    fuse_vae_embedding_in_latents: bool = False,
    **kwargs,
):
    for name, value in (("vace_context", vace_context), ("audio_embeds", audio_embeds), ("reference_latents", reference_latents),
                        ("tea_cache", tea_cache), ("pose_latents", pose_latents), ("face_pixel_values", face_pixel_values),
                        ("longcat_latents", longcat_latents), ("sliding_window_size", sliding_window_size),
                        ("control_camera_latents_input", control_camera_latents_input), ("vap_hidden_state", vap_hidden_state)):
        _reject(name, value)
    if motion_bucket_id is not None and motion_controller is not None:
        _reject("motion_controller", motion_controller)
    if y is not None and getattr(dit, "require_vae_embedding", False):
        _reject("y (VAE-concat image conditioning)", y)
    if clip_feature is not None and getattr(dit, "require_clip_embedding", False):
        _reject("clip_feature", clip_feature)
    if torch.is_grad_enabled() and any(p.requires_grad for p in dit.parameters()):
        raise NotImplementedError("fairygen_b200.model_fn_wan_video is the inference path (no autograd graph); for LoRA training use "
                                  "fairygen_b200.training.Stage2Trainer.model_fn, which back-propagates through the same kernels")
    if latents is None or timestep is None or context is None or dit is None:
        raise ValueError("model_fn_wan_video needs dit, latents, timestep and context")

    sp = None
    if use_unified_sequence_parallel:
        import torch.distributed as dist

        if dist.is_initialized() and dist.get_world_size() > 1:
            from .sp import SequenceParallel

            sp = getattr(dit, "_fairygen_b200_sp", None)
            if sp is None:
                sp = SequenceParallel()
                object.__setattr__(dit, "_fairygen_b200_sp", sp)
    eng = engine_for(dit, sp)
    if sp is not None:
        # a barrier of the exchange that gave up on a dead peer leaves its epoch in the arena (fgb_sp_barrier_status): look every
        # 32 forwards (one 4-byte read; a sync the host-side run-ahead absorbs) so that a lost rank ends the job with a message
        sp.forwards_since_check = getattr(sp, "forwards_since_check", 0) + 1
        if sp.forwards_since_check >= 32:
            sp.forwards_since_check = 0
            sp.check()

    # merged CFG (PIPE:785-803, 1240-1243): one latent, a batch of contexts -> one forward per context
    outs = []
    for b in range(context.shape[0]):
        lat_b = latents[b:b + 1] if latents.shape[0] == context.shape[0] else latents[0:1]
        outs.append(eng.forward(lat_b, timestep, context[b:b + 1], fuse_vae_embedding_in_latents))
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)


def install(pipe, fused_scheduler: bool = True, text_encoder: bool = False, vae: bool = False):
    """Point a reference ``WanVideoPipeline`` at the B200 path.  ``pipe.model_fn`` is the attribute the
    reference itself swaps behaviour through (wan_video.py:81); ``pipe.scheduler`` gets the fused
    flow-match step with identical ``set_timesteps`` / ``step`` semantics (flow_match.py:29-39, 132-154).
    text_encoder / vae = True additionally swap ``pipe.text_encoder`` (text_encoder.install) and route
    ``pipe.vae.decode`` / ``pipe.vae.encode`` (vae.install, vae_encode.install) through the kernels."""
    if text_encoder:
        from . import text_encoder as _te
        _te.install(pipe)
    if vae:
        from . import vae as _vae, vae_encode as _vae_encode
        _vae.install(pipe)
        _vae_encode.install(pipe)
    pipe.model_fn = model_fn_wan_video
    if fused_scheduler:
        from .scheduler import FlowMatchScheduler

        pipe.scheduler = FlowMatchScheduler("Wan")
    dit = getattr(pipe, "dit", None)
    if dit is not None and next(dit.parameters()).device.type == "cuda":
        engine_for(dit)        # pack now; a DiT that is still on the host is packed at its first model_fn call
    return pipe
