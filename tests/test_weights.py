"""Checkpoint detection / loading (fairygen_b200.weights) against the reference's conventions.  CPU only."""
import os

import pytest
import torch

from fairygen_b200 import TI2V_5B, WanDiTConfig, weights
from fairygen_b200.synthetic import param_shapes


def test_parameter_inventory_reproduces_the_reference_model_hash():
    """KNOWN ANSWER from the reference: configs/model_configs.py:291 lists md5 1f5ab7703c6fc803fdded85ff040c316 for the
    Wan2.2-TI2V-5B checkpoint (hash of its sorted key:shape list, core/loader/file.py:99-121). Our parameter inventory —
    the names and shapes the engine packs from and the synthetic weights are generated for — must hash to the same value,
    i.e. all 825 tensors of the model are accounted for with the right shapes."""
    shapes = param_shapes(TI2V_5B)
    assert len(shapes) == 825
    assert weights.keys_hash(shapes) == weights.TI2V_5B_HASH == "1f5ab7703c6fc803fdded85ff040c316"
    assert sum(int(torch.tensor(s).prod()) for s in shapes.values()) == 4_999_787_712     # parameter count


def test_detect_and_load_sharded_safetensors(tmp_path):
    import safetensors.torch as st
    cfg = WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    shapes = param_shapes(cfg)
    g = torch.Generator().manual_seed(0)
    sd = {k: torch.randn(s, generator=g).to(torch.bfloat16) for k, s in shapes.items()}
    names = sorted(sd)
    st.save_file({k: sd[k] for k in names[::2]}, str(tmp_path / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    st.save_file({k: sd[k] for k in names[1::2]}, str(tmp_path / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    pattern = os.path.join(str(tmp_path), "diffusion_pytorch_model*.safetensors")
    assert len(weights.expand(pattern)) == 2
    assert weights.file_shapes(pattern) == {k: list(v) for k, v in shapes.items()}
    assert weights.keys_hash(weights.file_shapes(pattern)) == weights.keys_hash(shapes)
    with pytest.raises(ValueError, match="Cannot detect the model type"):      # a tiny model is not a known checkpoint
        weights.detect(pattern)
    back = weights.load_state_dict(pattern)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        weights.expand(os.path.join(str(tmp_path), "nothing*.safetensors"))
