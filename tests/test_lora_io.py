"""LoRA checkpoint formats and the B1 + B2 merge (fairygen_b200.lora_io) against the reference's conventions
(animation/merge_weights.py:19-45, diffusion/logger.py:35-55, utils/lora/general.py:10-30).  CPU only."""
import os

import torch

from fairygen_b200 import lora_io
from oracle import wan_dit_oracle as o
from oracle import wan_train_oracle as t


def _stage1(cfg, dtype=torch.bfloat16):
    return {k: v.to(dtype) for k, v in o.make_lora(cfg, rank=32, seed=2).items()}


def test_merge_matches_reference_script(tmp_path):
    import safetensors.torch as st
    cfg = o.TINY
    s1 = _stage1(cfg)
    b2 = {k: v.to(torch.bfloat16) for k, v in t.make_b2(cfg).items()}
    s2 = lora_io.stage2_state_dict(b2)
    # stage-2 layout: every B2 twice (stripped + full parameter name), bf16
    assert len(s2) == 2 * len(b2) and all(v.dtype == torch.bfloat16 for v in s2.values())
    assert "blocks.0.self_attn.q.lora_B2.weight" in s2 and "pipe.dit.blocks.0.self_attn.q.lora_B2.weight" in s2
    p1, p2, pm = (str(tmp_path / n) for n in ("stage1.safetensors", "stage2.safetensors", "merged/out.safetensors"))
    st.save_file(s1, p1)
    st.save_file(s2, p2)
    merged = lora_io.merge_lora_weights(p1, p2, pm)
    assert os.path.exists(pm)
    back = st.load_file(pm)
    assert set(back) == set(s1)                                   # merged file keeps the stage-1 key names
    for k, v in s1.items():
        if "lora_A" in k:
            assert torch.equal(back[k], v)
        else:                                                     # B = B1 + B2 in the checkpoint dtype (merge_weights.py:40)
            mod = k[: -len(".lora_B.default.weight")]
            assert torch.equal(back[k], v + b2[mod]) and torch.equal(merged[k], back[k])
    # a missing B2 leaves B1 untouched (merge_weights.py:41-43)
    s2_missing = {k: v for k, v in s2.items() if "blocks.1.ffn.2" not in k}
    m2 = lora_io.merge_state_dicts(s1, s2_missing)
    assert torch.equal(m2["blocks.1.ffn.2.lora_B.default.weight"], s1["blocks.1.ffn.2.lora_B.default.weight"])
    # round trip of the stage-2 reader over both key spellings
    assert set(lora_io.load_stage2_checkpoint(p2)) == set(b2)
    only_full = {k: v for k, v in s2.items() if k.startswith("pipe.dit.")}
    got = lora_io.load_stage2_checkpoint(only_full)
    assert all(torch.equal(got[m], b2[m]) for m in b2)


def test_name_dict_follows_reference_key_normalisation(golden):
    cfg = o.TINY
    s1 = _stage1(cfg, torch.float32)
    assert lora_io.name_dict(s1) == o.lora_target_names(s1)        # the oracle's restatement is pinned by tests/golden/lora.npz
    alt = {"diffusion_model.blocks.0.self_attn.q.lora_up.weight": torch.zeros(4, 2), "diffusion_model.blocks.0.self_attn.q.lora_down.weight": torch.zeros(2, 4),
           "blocks.1.ffn.0.lora_B.weight": torch.zeros(4, 2), "blocks.1.ffn.0.lora_A.weight": torch.zeros(2, 4)}
    nd = lora_io.name_dict(alt)
    assert nd["blocks.0.self_attn.q"] == ("diffusion_model.blocks.0.self_attn.q.lora_up.weight", "diffusion_model.blocks.0.self_attn.q.lora_down.weight")
    assert nd["blocks.1.ffn.0"] == ("blocks.1.ffn.0.lora_B.weight", "blocks.1.ffn.0.lora_A.weight")
    assert lora_io.b2_key_for("blocks.3.cross_attn.v.lora_B.default.weight") == "blocks.3.cross_attn.v.lora_B2.weight"
