"""oracle/umt5_oracle.py against the REAL reference text encoder's outputs (tests/golden/umt5.npz, written by
oracle/make_golden_umt5.py in the build container): bucket function, per-layer position bias, T5 layer norm, gated-GELU
feed-forward, the encoder with and without mask, batch of two with different lengths, and the pipeline's zeroing."""
import os

import numpy as np
import pytest
import torch

from oracle import umt5_oracle as u

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "umt5.npz"))
CASES = {"short": (1, 40, (13,)), "pair": (2, 48, (48, 7)), "long": (1, 200, (170,))}


@pytest.fixture(scope="module")
def weights():
    return u.make_weights(u.TINY, seed=0)


def test_bucket_table_is_bit_exact():
    got = u.relative_position_bucket(torch.arange(-600, 601), 32, 128).numpy()
    assert np.array_equal(got, GOLD["buckets"])
    assert got.min() == 0 and got.max() == 31 and got[600] == 0          # rel = 0 -> bucket 0; both halves saturate
    assert got[600 + 128] == 31 and got[600 - 128] == 15


def test_position_bias_layer_norm_ffn(weights):
    cfg = u.TINY
    pb = u.position_bias(weights["blocks.1.pos_embedding.embedding.weight"], 20, 20, cfg.num_buckets, cfg.max_dist)
    assert np.array_equal(pb.numpy(), GOLD["pos_bias_l1"])
    x = torch.randn(3, 5, cfg.dim, generator=torch.Generator().manual_seed(8))
    np.testing.assert_allclose(u.t5_layer_norm(x, weights["blocks.1.norm1.weight"], cfg.eps).numpy(), GOLD["layer_norm"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(u.t5_ffn(weights, "blocks.1.ffn.", x).numpy(), GOLD["ffn"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", list(CASES))
def test_encoder_and_prompt_embedding(weights, name):
    cfg = u.TINY
    b, L, live = CASES[name]
    ids, mask = u.make_ids(cfg, b, L, live, seed=3)
    np.testing.assert_allclose(u.encoder_forward(weights, cfg, ids, mask).numpy(), GOLD[name], rtol=2e-5, atol=2e-5)
    emb = u.encode_prompt(weights, cfg, ids, mask).numpy()
    np.testing.assert_allclose(emb, GOLD[name + "_prompt"], rtol=2e-5, atol=2e-5)
    assert not emb[:, min(live):].any()           # the reference zeroes from the SHORTEST length on, in every sample


def test_encoder_without_mask(weights):
    ids, _ = u.make_ids(u.TINY, 1, 24, (24,), seed=5)
    np.testing.assert_allclose(u.encoder_forward(weights, u.TINY, ids).numpy(), GOLD["nomask"], rtol=2e-5, atol=2e-5)
