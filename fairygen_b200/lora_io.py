"""LoRA checkpoint formats of FairyGen's motion stage and the B1 + B2 merge.

Byte-compatible with the reference's files so stage-1 / stage-2 checkpoints flow between the reference and this
trainer in both directions:

  * stage 1 (``ModelLogger.save_model``, animation/diffsynth/diffusion/logger.py:35-55 after
    ``export_trainable_state_dict(remove_prefix="pipe.dit.")``): ``<module>.lora_A.default.weight`` [r, in],
    ``<module>.lora_B.default.weight`` [out, r], bf16 safetensors;
  * stage 2: every ``lora_B2`` twice — prefix-stripped ``<module>.lora_B2.weight`` (training_module.py:62-72) and under
    its full parameter name ``pipe.dit.<module>.lora_B2.weight`` added by hand (logger.py:43-48) — SURVEY §9 item 14;
  * merged (animation/merge_weights.py:19-45): A copied, ``B = B1 + B2`` in the checkpoint dtype, stage-1 key names;
    this is the file ``pipe.load_lora(pipe.dit, path, alpha=1)`` fuses at inference (inference.py:16-17).

``fuse_into_engine`` is the on-GPU equivalent of ``GeneralLoRALoader.fuse_lora_to_base_model`` (utils/lora/general.py:
44-62) for the packed engine weights: one ``fgb_lora_merge`` per adapted Linear instead of 300 ``load_state_dict`` round
trips through the module tree.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

A_SUFFIX, B_SUFFIX, B2_SUFFIX = ".lora_A.default.weight", ".lora_B.default.weight", ".lora_B2.weight"
FULL_PREFIX = "pipe.dit."


def _st():
    import safetensors.torch as st
    return st


def name_dict(lora_sd: Dict[str, torch.Tensor]) -> Dict[str, Tuple[str, str]]:
    """module name -> (B key, A key), the key normalisation of GeneralLoRALoader.get_name_dict (general.py:10-30):
    accepts ``lora_A/lora_B`` (with or without an adapter name) and ``lora_down/lora_up``, strips ``diffusion_model.``."""
    out: Dict[str, Tuple[str, str]] = {}
    for key in lora_sd:
        a_tag, b_tag = ("lora_down", "lora_up") if ".lora_up." in key else ("lora_A", "lora_B")
        parts = key.split(".")
        if b_tag not in parts:
            continue
        at = parts.index(b_tag)
        if len(parts) > at + 2:
            parts.pop(at + 1)
        parts.pop(at)
        if parts[0] == "diffusion_model":
            parts.pop(0)
        parts.pop(-1)
        out[".".join(parts)] = (key, key.replace(b_tag, a_tag))
    return out


def b2_key_for(stage1_b_key: str) -> str:
    """The stage-2 key merge_weights.py looks up for a stage-1 B key (merge_weights.py:34-37)."""
    if stage1_b_key.endswith(B_SUFFIX):
        return stage1_b_key.replace(B_SUFFIX, B2_SUFFIX)
    return stage1_b_key.replace("lora_B", "lora_B2").replace(".default", "")


def merge_state_dicts(stage1: Dict[str, torch.Tensor], stage2: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """merge_weights.py:28-43: A passes through, B = B1 + B2 (B1 alone, with a warning, when B2 is missing)."""
    merged: Dict[str, torch.Tensor] = {}
    for k, v in stage1.items():
        if "lora_A" in k:
            merged[k] = v
        elif "lora_B" in k:
            b2 = b2_key_for(k)
            if b2 in stage2:
                merged[k] = v + stage2[b2]
            else:
                print("Warning: Missing B2 key for", k, "-> expected", b2)
                merged[k] = v
    return merged


def merge_lora_weights(stage1_path: str, stage2_path: str, save_path: str) -> Dict[str, torch.Tensor]:
    """Drop-in for animation/merge_weights.py::merge_lora_weights (same inputs, same output file)."""
    import os
    st = _st()
    merged = merge_state_dicts(st.load_file(stage1_path), st.load_file(stage2_path))
    os.makedirs(os.path.dirname(os.path.abspath(save_path)), exist_ok=True)
    st.save_file(merged, save_path)
    return merged


def stage2_state_dict(b2: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """{module: B2} -> the stage-2 checkpoint layout (both key spellings, bf16, CPU)."""
    out: Dict[str, torch.Tensor] = {}
    for module, t in b2.items():
        v = t.detach().to("cpu").to(torch.bfloat16).contiguous()
        out[module + B2_SUFFIX] = v
        out[FULL_PREFIX + module + B2_SUFFIX] = v.clone()
    return out


def stage1_state_dict(trainer) -> Dict[str, torch.Tensor]:
    """A stage-1 trainer's adapters in the stage-1 checkpoint layout (``<module>.lora_{A,B}.default.weight``, bf16, CPU) —
    what ``export_trainable_state_dict(remove_prefix="pipe.dit.")`` writes (training_module.py:62-72, logger.py:41)."""
    out: Dict[str, torch.Tensor] = {}
    for module in trainer.targets:
        out[module + A_SUFFIX] = trainer.a1[module].detach().to("cpu").to(torch.bfloat16).contiguous()
        out[module + B_SUFFIX] = trainer.b2[module].detach().to("cpu").to(torch.bfloat16).contiguous()
    return out


def save_stage2_checkpoint(trainer, path: str) -> None:
    """What ModelLogger.save_model writes during stage-2 training, from a fairygen_b200 Stage2Trainer."""
    import os
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    _st().save_file(stage2_state_dict({t: trainer.b2[t] for t in trainer.targets}), path)


def load_stage2_checkpoint(path_or_sd) -> Dict[str, torch.Tensor]:
    """{module: B2} from a stage-2 checkpoint written by the reference or by save_stage2_checkpoint."""
    sd = _st().load_file(path_or_sd) if isinstance(path_or_sd, str) else path_or_sd
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if not k.endswith(B2_SUFFIX):
            continue
        module = k[: -len(B2_SUFFIX)]
        if module.startswith(FULL_PREFIX):
            module = module[len(FULL_PREFIX):]
        out.setdefault(module, v)
    return out


def fuse_into_engine(engine, lora_sd: Dict[str, torch.Tensor], alpha: float = 1.0, targets: Optional[Iterable[str]] = None,
                     _record: bool = True) -> int:
    """W <- W + alpha * B @ A on the engine's packed bf16 weights (fused QKV / cross-KV rows included), in place, on the GPU.
    Returns the number of fused Linears (the reference prints it, general.py:62)."""
    from . import ops

    names = name_dict(lora_sd)
    dev, d = engine.device, engine.cfg.dim
    slots = {}       # module name -> (block, weight attribute, first row, rows)
    for i, b in enumerate(engine.blocks):
        p = f"blocks.{i}."
        for j, proj in enumerate("qkv"):
            slots[p + "self_attn." + proj] = (b, "wqkv", j * d, d)
        slots[p + "self_attn.o"] = (b, "wo", 0, None)
        slots[p + "cross_attn.q"] = (b, "cwq", 0, None)
        for j, proj in enumerate("kv"):
            slots[p + "cross_attn." + proj] = (b, "cwkv", j * d, d)
        slots[p + "cross_attn.o"] = (b, "cwo", 0, None)
        slots[p + "ffn.0"] = (b, "w1", 0, None)
        slots[p + "ffn.2"] = (b, "w2", 0, None)
    wanted = set(targets) if targets is not None else None
    fused = 0
    for module, (kb, ka) in names.items():
        if module not in slots or (wanted is not None and module not in wanted):
            continue
        up = lora_sd[kb].detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        down = lora_sd[ka].detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        blk, attr, r0, nrows = slots[module]
        w = blk.own(attr)            # never the caller's parameter: weights that alias the module are copied on first write
        if nrows is not None:
            w = w[r0:r0 + nrows]
        ops.lora_merge(w, down, up, None, None, w, 1.0, float(alpha))   # in place: each output tile reads only itself
        fused += 1
    engine._ctx_cache.clear()          # cached cross-attention K/V were projected with the old weights
    engine._ctx_cache_order.clear()
    # remembered so that a re-pack from the weight container (model_fn.engine_for, after pipe.load_lora changed the module)
    # re-applies this adapter instead of silently dropping it; engine.load_state_dict() itself starts from a clean slate
    if _record and hasattr(engine, "fused_adapters"):
        engine.fused_adapters.append(({k: v.detach().clone() for k, v in lora_sd.items()}, float(alpha),
                                      None if targets is None else tuple(wanted)))
    return fused
