"""oracle/vae38_oracle.py against the REAL reference VAE38's outputs (tests/golden/vae38.npz, written by
oracle/make_golden_vae.py in the build container): chunked causal decode with the feature cache, single frame, tiled decode
with blending (even and ragged tilings), and the primitives (DupUp3D, unpatchify, attention block, temporal up-sampling
across chunks, residual block across chunks)."""
import os

import numpy as np
import pytest
import torch

from oracle import vae38_oracle as o

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "vae38.npz"))


def latents(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


@pytest.fixture(scope="module")
def weights():
    return o.make_weights(o.TINY, seed=0)


def close(a, b, tol=2e-5):
    np.testing.assert_allclose(a.numpy() if isinstance(a, torch.Tensor) else a, b, rtol=tol, atol=tol)


def test_plan_matches_the_full_size_decoder():
    assert o.VAE38.dims == [1024, 1024, 1024, 512, 256] and o.VAE38.upsampling_factor == 16
    assert [(up, t) for _, _, _, up, t in o.stage_plan(o.VAE38)] == [(True, True), (True, True), (True, False), (False, False)]
    assert o.count_cache_slots(o.VAE38) == o.count_cache_slots(o.TINY)


def test_model_decode_chunked(weights):
    with torch.no_grad():
        got = o.model_decode(weights, o.TINY, latents((1, 8, 3, 3, 4), 1))
    assert got.shape == (1, 3, 9, 48, 64)
    close(got, GOLD["model_decode"])


def test_single_frame_is_clamped(weights):
    with torch.no_grad():
        got = o.decode(weights, o.TINY, latents((1, 8, 1, 2, 2), 2) * 3)
    close(got, GOLD["single_frame"])
    assert got.min() >= -1 and got.max() <= 1 and (got.abs() == 1).any()


def test_tiled_decode(weights):
    with torch.no_grad():
        close(o.decode(weights, o.TINY, latents((1, 8, 2, 5, 5), 3), tiled=True, tile_size=(3, 3), tile_stride=(2, 2)), GOLD["tiled"])
        close(o.decode(weights, o.TINY, latents((1, 8, 1, 4, 7), 4), tiled=True, tile_size=(3, 4), tile_stride=(2, 3)), GOLD["tiled_ragged"])
    assert o.tile_tasks(44, 80, (30, 52), (15, 26)) == [(0, 30, 0, 52), (0, 30, 26, 78), (0, 30, 52, 104), (15, 45, 0, 52), (15, 45, 26, 78),
                                                        (15, 45, 52, 104)]          # 704x1280 with the pipeline's defaults


def test_primitives(weights):
    cfg = o.TINY
    x = latents((1, 32, 1, 3, 4), 5)
    assert np.array_equal(o.dup_up3d(x, 16, 2, 2, True).numpy(), GOLD["dup_t2_first"])
    assert np.array_equal(o.dup_up3d(x, 16, 2, 2, False).numpy(), GOLD["dup_t2"])
    assert np.array_equal(o.dup_up3d(x, 32, 1, 2, True).numpy(), GOLD["dup_t1"])
    assert np.array_equal(o.unpatchify(latents((1, 12, 2, 3, 5), 6)).numpy(), GOLD["unpatchify"])
    with torch.no_grad():
        close(o.attention_block(weights, "decoder.middle.1.", latents((1, cfg.dims[0], 2, 3, 4), 7)), GOLD["attention"])
        cache = [None]
        for i in range(3):          # chunk 0 skips the temporal doubling, chunk 1 sees zeros, chunk 2 the cached frame
            got = o.resample_up(weights, "decoder.upsamples.0.upsamples.3.", latents((1, cfg.dims[1], 1, 2, 3), 10 + i), True, cache, [0])
            close(got, GOLD[f"resample3d_c{i}"])
        cache = [None, None]
        for i in range(3):
            got = o.residual_block(weights, "decoder.upsamples.2.upsamples.0.", latents((1, cfg.dims[2], 1 if i < 2 else 2, 3, 3), 20 + i),
                                   cache, [0])
            close(got, GOLD[f"resblock_c{i}"])


def test_host_mirror_agrees_with_the_oracle_on_keys_shapes_and_tiling():
    """fairygen_b200.vae (product, CUDA only) and the oracle list the same state-dict keys / shapes and the same tile windows."""
    from fairygen_b200 import vae
    assert vae.param_shapes(vae.VAE38) == o.param_shapes(o.VAE38)
    assert vae.param_shapes(vae.VAE38Config(z_dim=8, dec_dim=16)) == o.param_shapes(o.TINY)
    for args in [(44, 80, (30, 52), (15, 26)), (5, 5, (3, 3), (2, 2)), (4, 7, (3, 4), (2, 3)), (30, 52, (34, 34), (18, 16))]:
        assert vae.tile_tasks(*args) == o.tile_tasks(*args)
    assert vae.MEAN38 == o.MEAN38 and vae.STD38 == o.STD38


def test_window_assignment_covers_every_window_once_and_balances():
    from fairygen_b200 import vae
    tasks = vae.tile_tasks(44, 80, (30, 52), (15, 26))                      # the 6 windows of the 704x1280 decode
    for world in (1, 2, 4, 6, 8):
        parts = vae.assign_windows(tasks, 44, 80, world)
        assert len(parts) == world and sorted(t for p in parts for t in p) == sorted(tasks)
        area = lambda t: (min(t[1], 44) - t[0]) * (min(t[3], 80) - t[2])  # noqa: E731
        loads = [sum(area(t) for t in p) for p in parts]
        assert max(loads) <= {1: 7788, 2: 3908, 4: 2348, 6: 1560, 8: 1560}[world]
    assert vae.assign_windows(tasks, 44, 80, 2) == vae.assign_windows(tasks, 44, 80, 2)      # deterministic: every rank agrees


# ---------------------------------------------------------------------------------------------------------------
# encoder (first-frame conditioning)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def enc_weights():
    return o.make_enc_weights(o.TINY, seed=0)


def test_encoder_plan_and_primitives():
    assert o.VAE38.enc_dims == [160, 160, 320, 640, 640]
    assert [(d, t) for _, _, _, d, t in o.enc_stage_plan(o.VAE38)] == [(True, False), (True, True), (True, True), (False, False)]
    xa = latents((1, 16, 3, 4, 6), 33)
    assert np.array_equal(o.avg_down3d(xa, 32, 2, 2).numpy(), GOLD["avg_down_t2"])
    assert np.array_equal(o.avg_down3d(xa, 16, 1, 2).numpy(), GOLD["avg_down_t1"])
    x = latents((1, 3, 2, 6, 10), 34)
    assert np.array_equal(o.patchify(x).numpy(), GOLD["patchify"])
    assert torch.equal(o.unpatchify(o.patchify(x)), x)


def test_encode_image_clip_and_tiled(enc_weights):
    cfg = o.TINY
    with torch.no_grad():
        close(o.encode(enc_weights, cfg, [torch.tanh(latents((3, 1, 32, 48), 30))]), GOLD["encode_image"])
        close(o.encode(enc_weights, cfg, [torch.tanh(latents((3, 9, 32, 32), 31))]), GOLD["encode_clip"])      # chunks of 1, 4, 4 frames
        close(o.encode(enc_weights, cfg, [torch.tanh(latents((3, 1, 80, 96), 32))], tiled=True, tile_size=(3, 4), tile_stride=(2, 2)),
              GOLD["encode_tiled"])
