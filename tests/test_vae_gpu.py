"""VAE38 decoder on the B200 against the pinned oracle (oracle/vae38_oracle.py, itself checked against the real reference VAE
in tests/test_vae_oracle.py): the tap-GEMM convolution and the memory-bound kernels alone, every intermediate grid of the
reduced-width decoder, the reference's stored outputs (chunked decode, single frame, tiled decode with even and ragged
tilings), and the full-width decoder on a small window."""
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
    from fairygen_b200 import ops, vae
    from oracle import vae38_oracle as o
    torch.cuda.set_device(0)
    ops.context(torch.device("cuda", 0))
    return ops, vae, o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


def to_grid(x, cp):
    """[C, T, H, W] -> zero-bordered channels-last rows [T*(H+2)*(W+2), cp]."""
    C, T, H, W = x.shape
    g = torch.zeros(T, H + 2, W + 2, cp, dtype=BF, device=x.device)
    g[:, 1:-1, 1:-1, :C] = x.permute(1, 2, 3, 0)
    return g.view(-1, cp)


def from_grid(rows, C, T, H, W):
    return rows.float().view(T, H + 2, W + 2, -1)[:, 1:-1, 1:-1, :C].permute(3, 0, 1, 2)


@pytest.mark.parametrize("cin,cout,T,H,W", [(64, 64, 1, 3, 4), (128, 320, 2, 5, 7), (256, 64, 4, 9, 6)])
def test_causal_conv3d_as_tap_gemm(env, cin, cout, T, H, W):
    """3x3x3 CausalConv3d with two cached frames in front == F.conv3d on (cache | x) with zero padding in h, w."""
    ops, vae, o = env
    w = rnd(cout, cin, 3, 3, 3, seed=1, scale=(27 * cin) ** -0.5)
    b = rnd(cout, seed=2, scale=0.1)
    x = rnd(cin, 2 + T, H, W, seed=3)                                   # frames 0, 1 = the cache
    conv = vae._Conv(w, b, "cuda")
    P = (H + 2) * (W + 2)
    out = torch.full((T * P, conv.n), float("nan"), dtype=BF, device="cuda")
    ops.conv_taps(to_grid(x, conv.cin_p), 2 * P, conv.w, conv.b, out, conv.offsets(H + 2, W + 2), (H + 2, W + 2))
    ops.sync_check()
    want = F.conv3d(F.pad(x.float()[None], (1, 1, 1, 1, 0, 0)), w.float(), b.float())[0]
    assert rel_l2(from_grid(out, cout, T, H, W), want) < 4e-3
    g = out.float().view(T, H + 2, W + 2, -1)
    assert g[:, 0].abs().max() == 0 and g[:, -1].abs().max() == 0 and g[:, :, 0].abs().max() == 0 and g[:, :, -1].abs().max() == 0
    assert g[..., cout:].abs().max() == 0 if conv.n > cout else True
    # residual epilogue: out += conv
    base = to_grid(rnd(cout, T, H, W, seed=4), conv.n)
    acc = base.clone()
    ops.conv_taps(to_grid(x, conv.cin_p), 2 * P, conv.w, conv.b, acc, conv.offsets(H + 2, W + 2), (H + 2, W + 2), ops.EPI_RESIDUAL)
    ops.sync_check()
    assert rel_l2(from_grid(acc, cout, T, H, W), want + from_grid(base, cout, T, H, W)) < 4e-3


def test_norm_upsample_dup_softmax_unpatchify(env):
    ops, vae, o = env
    C, T, H, W = 96, 2, 3, 5
    x = rnd(C, T, H, W, seed=1, scale=2.0)
    gamma = torch.zeros(128, dtype=BF, device="cuda")
    gamma[:C] = 1 + rnd(C, seed=2, scale=0.2)
    gx = to_grid(x, 128)
    y = torch.empty_like(gx)
    ops.vae_norm_silu(gx, y, C, gamma, silu=True)
    want = F.silu(o.rms_norm(x.float()[None], gamma[:C].float().view(C, 1, 1, 1)))[0]
    assert rel_l2(from_grid(y, C, T, H, W), want) < 4e-3
    assert y.float().view(T, H + 2, W + 2, 128)[:, 0].abs().max() == 0
    # nearest-exact x2, plain and with the frame interleave of the temporal up-sampling
    up = torch.zeros(T * (2 * H + 2) * (2 * W + 2), 128, dtype=BF, device="cuda")
    ops.vae_upsample2x(gx, up, 128, T, H, W, halves=1)
    want = F.interpolate(x.float().permute(1, 0, 2, 3), scale_factor=(2.0, 2.0), mode="nearest-exact").permute(1, 0, 2, 3)
    assert torch.equal(from_grid(up, C, T, 2 * H, 2 * W), want)
    both = rnd(2 * 64, T, H, W, seed=3)                                                  # channels [frame-a | frame-b]
    up2 = torch.zeros(2 * T * (2 * H + 2) * (2 * W + 2), 64, dtype=BF, device="cuda")
    ops.vae_upsample2x(to_grid(both, 128), up2, 64, 2 * T, H, W, halves=2)
    inter = torch.stack((both[:64], both[64:]), 2).reshape(64, 2 * T, H, W).float()     # VAE:152-156
    want = F.interpolate(inter.permute(1, 0, 2, 3), scale_factor=(2.0, 2.0), mode="nearest-exact").permute(1, 0, 2, 3)
    assert torch.equal(from_grid(up2, 64, 2 * T, 2 * H, 2 * W), want)
    # DupUp3D shortcut added onto a grid
    for cin, cout, ft, first in [(128, 64, 2, True), (128, 64, 2, False), (64, 64, 1, True), (64, 32, 1, False)]:
        xs = rnd(cin, T, H, W, seed=5)
        t_out = T * ft - (ft - 1 if first else 0)
        main = rnd(cout, t_out, 2 * H, 2 * W, seed=6)
        gm = to_grid(main, 64)
        ops.vae_dup_up_add(to_grid(xs, 128), gm, cin, cout, ft, first, t_out, H, W)
        want = main.float() + o.dup_up3d(xs.float()[None], cout, ft, 2, first)[0]
        assert rel_l2(from_grid(gm, cout, t_out, 2 * H, 2 * W), want) < 3e-3, (cin, cout, ft, first)
    # softmax over the interior positions of one frame
    gh, gw = H + 2, W + 2
    P, n8 = gh * gw, -(-gh * gw // 8) * 8
    s = rnd(P, n8, seed=7, scale=3.0)
    s0 = s.float().clone()
    ops.vae_attn_softmax(s, n8, gh, gw, 0.5)
    live = torch.zeros(n8, dtype=torch.bool, device="cuda")
    yy, xx = torch.arange(gh, device="cuda")[:, None], torch.arange(gw, device="cuda")[None, :]
    live[:P] = ((yy > 0) & (yy < gh - 1) & (xx > 0) & (xx < gw - 1)).reshape(-1)
    want = torch.softmax((s0 * 0.5).masked_fill(~live, float("-inf")), dim=-1)
    assert rel_l2(s, want) < 4e-3 and s[:, ~live].abs().max() == 0
    # un-patchify: direct (clamped) and blended
    head = rnd(12, T, H, W, seed=8, scale=0.8)
    gh_rows = to_grid(head, 64)
    values = torch.zeros(3, T + 1, 2 * H + 2, 2 * W, dtype=torch.float32, device="cuda")
    ops.vae_unpatchify(gh_rows, T, H, W, values, None, 1, 2, 0)
    want = o.unpatchify(head.float()[None])[0].clamp(-1, 1)
    assert torch.equal(values[:, 1:, 2:], want) and values[:, 0].abs().max() == 0
    values.zero_()
    weight = torch.zeros(T + 1, 2 * H + 2, 2 * W, dtype=torch.float32, device="cuda")
    ops.vae_unpatchify(gh_rows, T, H, W, values, weight, 1, 2, 0, bounds=(False, True, True, False), border=(3, 4))
    mh, mw = o.build_1d_mask(2 * H, False, True, 3).cuda(), o.build_1d_mask(2 * W, True, False, 4).cuda()
    mask = torch.minimum(mh[:, None], mw[None, :])
    assert torch.allclose(values[:, 1:, 2:], o.unpatchify(head.float()[None])[0] * mask, atol=1e-6)
    assert torch.allclose(weight[1:, 2:], mask.expand(T, -1, -1), atol=1e-7)
    ops.vae_blend_finish(values, weight.clamp_(min=1e-3))
    ops.sync_check()
    assert values.abs().max() <= 1


def _tiny(env):
    ops, vae, o = env
    ocfg = o.TINY
    w = o.make_weights(ocfg, seed=0)
    dec = vae.VAE38Decoder(vae.VAE38Config(z_dim=ocfg.z_dim, dec_dim=ocfg.dec_dim), "cuda")
    dec.load_state_dict({"model." + k: v for k, v in w.items()})         # the WanVideoVAE38 key layout
    return ocfg, w, dec


def latents(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def test_every_intermediate_grid_matches_the_oracle(env):
    """3 latent frames (chunks of 1, 4, 4 output frames: feature cache, 'Rep' first chunk, both temporal up-samplings):
    each traced grid against the fp32 oracle on the same bf16-rounded weights; grid borders and padding channels stay zero."""
    ops, vae, o = env
    ocfg, w, dec = _tiny(env)
    z = latents((1, ocfg.z_dim, 3, 3, 4), 1).to(BF)
    got = []

    def trace(name, rows, T, h, w_, C):
        g = rows.float().view(T, h + 2, w_ + 2, -1)
        edge = g.clone()
        edge[:, 1:-1, 1:-1, :C] = 0
        got.append((name, g[:, 1:-1, 1:-1, :C].permute(3, 0, 1, 2).contiguous(), float(edge.abs().max())))

    dec.trace = trace
    out = dec.decode(z)
    ops.sync_check()
    want = []
    w16 = {k: v.to(BF).float().cuda() for k, v in w.items()}
    with torch.no_grad():
        ref = o.model_decode(w16, ocfg, z.float().cuda(), trace=lambda n, t: want.append((n, t[0])))
        chain = o.model_decode({k: v.to(BF) for k, v in w16.items()}, ocfg, z.cuda())      # the reference's own bf16 op chain
    assert len(got) == len(want) == 3 * 23
    for (n1, a, edge), (n2, b) in zip(got, want):
        assert n1 == n2 and tuple(a.shape) == tuple(b.shape), (n1, n2, a.shape, b.shape)
        assert edge == 0, (n1, edge)
        assert rel_l2(a, b) < 2e-2, (n1, rel_l2(a, b))
    ours, theirs = rel_l2(out[0], ref[0].clamp(-1, 1)), rel_l2(chain[0].float().clamp(-1, 1), ref[0].clamp(-1, 1))
    print(f"VAE tiny decode vs fp32 oracle: ours {ours:.3e}, torch bf16 chain {theirs:.3e}, ours vs chain "
          f"{rel_l2(out[0], chain[0].float().clamp(-1, 1)):.3e}")
    assert ours < 2e-2 and ours < 1.5 * theirs + 2e-3        # no worse than the reference's bf16 arithmetic itself
    gold = torch.from_numpy(np.load(GOLD)["model_decode"]).cuda().clamp(-1, 1)
    assert rel_l2(out, gold) < 2.5e-2


def test_reference_goldens_single_and_tiled(env):
    ops, vae, o = env
    ocfg, w, dec = _tiny(env)
    gold = np.load(GOLD)
    single = dec.decode((latents((1, ocfg.z_dim, 1, 2, 2), 2) * 3).to(BF))
    assert single.dtype == BF and rel_l2(single, torch.from_numpy(gold["single_frame"]).cuda()) < 2.5e-2
    assert single.float().abs().max() <= 1
    tiled = dec.decode(latents((1, ocfg.z_dim, 2, 5, 5), 3).to(BF), tiled=True, tile_size=(3, 3), tile_stride=(2, 2))
    assert rel_l2(tiled, torch.from_numpy(gold["tiled"]).cuda()) < 2.5e-2
    ragged = dec.decode(latents((1, ocfg.z_dim, 1, 4, 7), 4).to(BF), tiled=True, tile_size=(3, 4), tile_stride=(2, 3))
    ops.sync_check()
    assert rel_l2(ragged, torch.from_numpy(gold["tiled_ragged"]).cuda()) < 2.5e-2
    # a list of latents of different sizes, as WanVideoVAE.decode accepts (VAE:1236-1246)
    assert dec.decode([latents((ocfg.z_dim, 1, 2, 2), 2).to(BF) * 3])[0].shape == (3, 1, 32, 32)


def test_full_width_decoder_on_a_small_window(env):
    """dec_dim 256 (1024 / 512 / 256 channels, 48 latent channels): two latent frames of a 4 x 6 window — every convolution
    at its production channel counts (multi-block K per tap, 1-4 N tiles), attention with head width 1024."""
    ops, vae, o = env
    w = o.make_weights(o.VAE38, seed=1)
    dec = vae.VAE38Decoder(vae.VAE38, "cuda")
    dec.load_state_dict(w)
    z = latents((1, 48, 2, 4, 6), 5).to(BF)
    out = dec.decode(z)
    ops.sync_check()
    w16 = {k: v.to(BF).float().cuda() for k, v in w.items()}
    with torch.no_grad():
        ref = o.model_decode(w16, o.VAE38, z.float().cuda()).clamp(-1, 1)
    assert out.shape == (1, 3, 5, 64, 96) and torch.isfinite(out.float()).all()
    print(f"VAE full-width window vs fp32 oracle: {rel_l2(out, ref):.3e}")
    assert rel_l2(out, ref) < 2e-2


def test_vae_rejects_bad_arguments(env):
    ops, vae, o = env
    ocfg, w, dec = _tiny(env)
    with pytest.raises(RuntimeError):
        vae.VAE38Decoder(dec.cfg, "cuda").decode(torch.zeros(1, 8, 1, 2, 2))
    with pytest.raises(ValueError):
        dec.decode(torch.zeros(1, 7, 1, 2, 2))
    with pytest.raises(ValueError):
        dec.decode(torch.zeros(1, 8, 1, 4, 4), tiled=True, tile_size=(2, 2), tile_stride=(2, 2))
    with pytest.raises(KeyError):
        vae.VAE38Decoder(dec.cfg, "cuda").load_state_dict({k: v for k, v in w.items() if k != "decoder.head.2.bias"})
    with pytest.raises(RuntimeError):
        ops.conv_taps(rnd(32, 48), 0, rnd(64, 48), None, torch.empty(32, 64, dtype=BF, device="cuda"), [0])   # cin % 64
