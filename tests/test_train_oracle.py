"""The training oracle (oracle/wan_train_oracle.py) against tests/golden/train.npz — loss, prediction and every
lora_B2 gradient of one stage-2 fine-tune step computed by the REAL reference DiT under autograd
(oracle/make_golden_train.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import wan_dit_oracle as o
from oracle import wan_train_oracle as t


@pytest.mark.parametrize("tag,shape,text_len,timestep_id", [("a", (1, 48, 3, 8, 8), 32, 500), ("b", (1, 48, 2, 6, 10), 24, 37)])
def test_training_step_matches_reference(golden, tag, shape, text_len, timestep_id):
    g = golden("train")
    cfg = o.TINY
    w, lora = o.make_weights(cfg, seed=0), o.make_lora(cfg, rank=32, seed=2)
    b2, masks = t.make_b2(cfg), t.make_masks(cfg)
    x0, _, ctx, _ = o.make_inputs(cfg, shape, text_len=text_len, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    loss, pred, grads = t.loss_and_grads(w, cfg, lora, b2, masks, x0, noise, timestep_id, ctx)
    assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 1e-6 * max(1.0, abs(float(g[f"{tag}_loss"])))
    np.testing.assert_allclose(pred.numpy(), g[f"{tag}_pred"], rtol=1e-5, atol=1e-6)
    assert len(grads) == 20
    for name, gr in grads.items():
        ref = g[f"{tag}_grad.{name}"]
        assert np.abs(ref).max() > 0, name            # every adapted Linear (incl. cross-attention k/v) gets a gradient
        np.testing.assert_allclose(gr.numpy(), ref, rtol=2e-4, atol=1e-8 + 2e-5 * np.abs(ref).max(), err_msg=name)
        # the weight-dropout mask zeroes exactly the dropped entries
        assert np.all(ref[masks[name].numpy() == 0] == 0)


def test_training_schedule_and_targets():
    sig, ts, wts = t.training_schedule()
    assert len(sig) == 1000 and float(ts[0]) == 1000.0 and abs(float(wts.sum()) - 1000.0) < 1e-2
    cfg = o.TINY
    assert len(list(t.lora_targets(cfg))) == 10 * cfg.num_layers
    m = t.make_masks(cfg)
    frac = np.mean([float(v.float().mean()) for v in m.values()])
    assert 0.45 < frac < 0.55


def test_stage1_training_step_matches_reference(golden):
    """Stage-1 (identity) LoRA: A and B trainable, weight dropout 0.8 (training_module.py:200-264) — loss, prediction and
    the stored A / B gradients of the REAL reference DiT under autograd."""
    g = golden("train")
    cfg = o.TINY
    w, lora = o.make_weights(cfg, seed=0), o.make_lora(cfg, rank=32, seed=2)
    masks = t.make_masks_stage1(cfg)
    shape = (1, 48, 3, 8, 8)
    x0, _, ctx, _ = o.make_inputs(cfg, shape, text_len=32, live_text=8)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9))
    loss, pred, grads = t.loss_and_grads_stage1(w, cfg, lora, masks, x0, noise, 500, ctx)
    assert abs(float(loss) - float(g["s1_loss"])) <= 1e-6 * abs(float(g["s1_loss"]))
    np.testing.assert_allclose(pred.numpy(), g["s1_pred"], rtol=1e-5, atol=1e-6)
    stored = [k[len("s1_gradA."):] for k in g if k.startswith("s1_gradA.")]
    assert len(stored) == 6
    for name in stored:
        for tag, key in (("A", f"{name}.lora_A.default.weight"), ("B", f"{name}.lora_B.default.weight")):
            ref = g[f"s1_grad{tag}.{name}"]
            assert np.abs(ref).max() > 0, (name, tag)
            np.testing.assert_allclose(grads[key].numpy(), ref, rtol=2e-4, atol=1e-8 + 2e-5 * np.abs(ref).max(), err_msg=f"{name} {tag}")
    frac = np.mean([float(v.float().mean()) for v in masks.values()])
    assert 0.17 < frac < 0.23          # keep probability 0.2
