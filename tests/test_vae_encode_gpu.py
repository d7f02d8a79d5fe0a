"""VAE38 encoder on the B200: the encoder-side kernels (patchify, space-to-depth stride-2 convolution, AvgDown3D, latent output)
against plain torch, and the encoder (image, 9-frame clip, tiled) against the pinned oracle and the reference's stored outputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vae38.npz")


@pytest.fixture(scope="module")
def env():
    from fairygen_b200 import ops, vae, vae_encode
    from oracle import vae38_oracle as o
    torch.cuda.set_device(0)
    ops.context(torch.device("cuda", 0))
    return ops, vae, vae_encode, o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


def to_grid(x, cp):
    C, T, H, W = x.shape
    g = torch.zeros(T, H + 2, W + 2, cp, dtype=BF, device=x.device)
    g[:, 1:-1, 1:-1, :C] = x.permute(1, 2, 3, 0)
    return g.view(-1, cp)


def from_grid(rows, C, T, H, W):
    return rows.float().view(T, H + 2, W + 2, -1)[:, 1:-1, 1:-1, :C].permute(3, 0, 1, 2)


def test_patchify_space_to_depth_stride2_conv(env):
    ops, vae, ve, o = env
    video = rnd(3, 2, 8, 12, seed=1)
    grid = torch.zeros(2 * 6 * 8, 64, dtype=BF, device="cuda")
    ops.vae_patchify_rows(video, grid, 64)
    assert torch.equal(from_grid(grid, 12, 2, 4, 6), o.patchify(video.float()[None])[0])
    # stride-2 3x3 convolution behind ZeroPad2d((0, 1, 0, 1)) == space-to-depth + 2x2-tap GEMM
    cin, cout, T, H, W = 96, 160, 2, 6, 10
    x = rnd(cin, T, H, W, seed=2)
    w = rnd(cout, cin, 3, 3, seed=3, scale=(9 * cin) ** -0.5)
    b = rnd(cout, seed=4, scale=0.1)
    conv = ve._ConvS2(w, b, "cuda")
    s2d = torch.zeros(T * (H // 2 + 2) * (W // 2 + 2), 4 * 128, dtype=BF, device="cuda")
    ops.vae_space_to_depth(to_grid(x, 128), s2d, 128, T, H, W)
    out = torch.empty(T * (H // 2 + 2) * (W // 2 + 2), conv.n, dtype=BF, device="cuda")
    ops.conv_taps(s2d, 0, conv.w, conv.b, out, conv.offsets(H // 2 + 2, W // 2 + 2), (H // 2 + 2, W // 2 + 2))
    ops.sync_check()
    want = F.conv2d(F.pad(x.float().permute(1, 0, 2, 3), (0, 1, 0, 1)), w.float(), b.float(), stride=2).permute(1, 0, 2, 3)
    assert rel_l2(from_grid(out, cout, T, H // 2, W // 2), want) < 4e-3


def test_avg_down_and_latent_out(env):
    ops, vae, ve, o = env
    for cin, cout, ft, fs, T in [(16, 32, 2, 2, 4), (16, 32, 2, 2, 1), (16, 16, 1, 2, 3), (64, 64, 1, 1, 2)]:
        H, W = 4 * fs, 3 * fs
        x = rnd(cin, T, H, W, seed=5)
        pad = (ft - T % ft) % ft
        t_out = (T + pad) // ft
        main = rnd(cout, t_out, H // fs, W // fs, seed=6)
        gm = to_grid(main, 64)
        ops.vae_avg_down_add(to_grid(x, 64), gm, cin, cout, ft, fs, pad, t_out, H // fs, W // fs)
        want = main.float() + o.avg_down3d(x.float()[None], cout, ft, fs)[0]
        assert rel_l2(from_grid(gm, cout, t_out, H // fs, W // fs), want) < 4e-3, (cin, cout, ft, fs, T)
    z, T, h, w = 8, 2, 3, 4
    mu = rnd(2 * z, T, h, w, seed=7)
    mean, inv_std = o.latent_scale(o.TINY)
    mean, inv_std = mean.cuda(), inv_std.cuda()
    values = torch.zeros(z, T, h + 1, w + 2, dtype=torch.float32, device="cuda")
    ops.vae_latent_out(to_grid(mu, 64), T, h, w, mean, inv_std, values, None, 0, 1, 2)
    want = (mu[:z].float() - mean.view(-1, 1, 1, 1)) * inv_std.view(-1, 1, 1, 1)
    ops.sync_check()
    assert torch.allclose(values[:, :, 1:, 2:], want, atol=1e-6) and values[:, :, 0].abs().max() == 0


def _encoder(env):
    ops, vae, ve, o = env
    w = o.make_enc_weights(o.TINY, seed=0)
    enc = ve.VAE38Encoder(vae.VAE38Config(z_dim=o.TINY.z_dim, dec_dim=o.TINY.dec_dim), "cuda", enc_dim=o.TINY.enc_dim)
    enc.load_state_dict({"model." + k: v for k, v in w.items()})
    return enc, {k: v.to(BF).float().cuda() for k, v in w.items()}


def latents(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("shape,seed,gold", [((3, 1, 32, 48), 30, "encode_image"), ((3, 9, 32, 32), 31, "encode_clip")])
def test_encoder_vs_oracle_and_reference_golden(env, shape, seed, gold):
    ops, vae, ve, o = env
    enc, w16 = _encoder(env)
    video = torch.tanh(latents(shape, seed)).to(BF)
    got = enc.encode([video])
    ops.sync_check()
    with torch.no_grad():
        want = o.encode(w16, o.TINY, [video.float().cuda()])
    assert rel_l2(got, want) < 2e-2
    assert rel_l2(got, torch.from_numpy(np.load(GOLD)[gold]).cuda()) < 3e-2


def test_tiled_encode_vs_reference_golden(env):
    ops, vae, ve, o = env
    enc, _ = _encoder(env)
    got = enc.encode([torch.tanh(latents((3, 1, 80, 96), 32)).to(BF)], tiled=True, tile_size=(3, 4), tile_stride=(2, 2))
    ops.sync_check()
    assert rel_l2(got, torch.from_numpy(np.load(GOLD)["encode_tiled"]).cuda()) < 3e-2
