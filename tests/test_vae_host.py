"""Host logic of the VAE38 encoder (fairygen_b200/vae_encode.py) and decoder (vae.py) on the CPU: the kernels are replaced by plain-torch statements of
what each C entry point computes (the same contracts include/fairygen_b200.h states), and the orchestration — chunking, feature
cache, tap offsets, space-to-depth stride-2 convolution, per-frame temporal convolution, AvgDown3D arguments, tiling — must then
reproduce the pinned oracle.  The kernels themselves are covered by tests/test_vae_encode_gpu.py."""
import numpy as np
import pytest
import torch

from oracle import vae38_oracle as o

BF = torch.bfloat16


def _emulated_ops(monkeypatch):
    from fairygen_b200 import ops

    def conv_taps(x, a_row0, w, bias, out, tap_offsets, grid_hw=(0, 0), epilogue=ops.EPI_BIAS):
        m, n = out.shape
        cin = x.shape[1]
        acc = torch.zeros(m, n)
        base = torch.arange(m) + a_row0
        for t, off in enumerate(tap_offsets):
            idx = base + off
            ok = (idx >= 0) & (idx < x.shape[0])
            rows = torch.zeros(m, cin)
            rows[ok] = x[idx[ok]].float()
            acc += rows @ w[:, t * cin:(t + 1) * cin].float().T
        y = (acc + (0 if bias is None else bias.float())).to(BF).float()
        if epilogue == ops.EPI_RESIDUAL:
            y = out.float() + y
        if grid_hw[1] > 0:
            pos = torch.arange(m) % (grid_hw[0] * grid_hw[1])
            gy, gx = pos // grid_hw[1], pos % grid_hw[1]
            y[(gy == 0) | (gy == grid_hw[0] - 1) | (gx == 0) | (gx == grid_hw[1] - 1)] = 0
        out.copy_(y.to(BF))
        return out

    def vae_norm_silu(x, out, channels, gamma, silu=True):
        xf = x.float()
        y = (xf / xf.norm(dim=1, keepdim=True).clamp_min(1e-12) * channels ** 0.5 * gamma.float()).to(BF).float()
        out.copy_((torch.nn.functional.silu(y) if silu else y).to(BF))

    def vae_patchify_rows(video, grid, cp):
        _, T, H, W = video.shape
        g = grid.view(T, H // 2 + 2, W // 2 + 2, cp)
        for c in range(3):
            for r in range(2):
                for q in range(2):
                    g[:, 1:-1, 1:-1, (c * 2 + r) * 2 + q] = video[c, :, q::2, r::2]

    def vae_space_to_depth(src, dst, cp, frames, h, w):
        s = src[:frames * (h + 2) * (w + 2)].view(frames, h + 2, w + 2, cp)
        d = dst[:frames * (h // 2 + 2) * (w // 2 + 2)].view(frames, h // 2 + 2, w // 2 + 2, 4 * cp)
        for py in range(2):
            for px in range(2):
                d[:, 1:-1, 1:-1, (py * 2 + px) * cp:(py * 2 + px + 1) * cp] = s[:, 1 + py:1 + h:2, 1 + px:1 + w:2]

    def vae_avg_down_add(x, main, cin, cout, factor_t, factor_s, pad_front, frames_out, h_out, w_out):
        hin, win = h_out * factor_s, w_out * factor_s
        t_in = frames_out * factor_t - pad_front
        xv = x[:t_in * (hin + 2) * (win + 2)].view(t_in, hin + 2, win + 2, -1)[:, 1:-1, 1:-1, :cin].float().permute(3, 0, 1, 2)[None]
        add = o.avg_down3d(xv, cout, factor_t, factor_s)[0]                      # [cout, frames_out, h_out, w_out]
        m = main[:frames_out * (h_out + 2) * (w_out + 2)].view(frames_out, h_out + 2, w_out + 2, -1)
        m[:, 1:-1, 1:-1, :cout] = (m[:, 1:-1, 1:-1, :cout].float() + add.to(BF).float().permute(1, 2, 3, 0)).to(BF)

    def gemm(a, w, bias, out, *args, **kw):
        out.copy_((a.float() @ w.float().T).to(BF))
        return out

    def gemm_dgrad(dy, w, dx, *args, **kw):
        dx.copy_((dy.float() @ w.float()).to(BF))
        return dx

    def vae_attn_softmax(scores, n_cols, gh, gw, scale):
        col = torch.arange(n_cols)
        gy, gx = col // gw, col % gw
        live = (col < gh * gw) & (gy > 0) & (gy < gh - 1) & (gx > 0) & (gx < gw - 1)
        s = (scores[:, :n_cols].float() * scale).masked_fill(~live, float("-inf"))
        scores[:, :n_cols] = torch.softmax(s, dim=1).to(BF)

    def ramp(n, first, last, border):
        return o.build_1d_mask(n, first, last, border)

    def vae_latent_out(grid, frames, h, w, mean, inv_std, values, weight, t0, y0, x0, bounds=(True,) * 4, border=(1, 1)):
        z = values.shape[0]
        g = grid[:frames * (h + 2) * (w + 2)].view(frames, h + 2, w + 2, -1)[:, 1:-1, 1:-1, :z].float()
        v = ((g - mean) * inv_std).permute(3, 0, 1, 2)
        if weight is None:
            values[:, t0:t0 + frames, y0:y0 + h, x0:x0 + w] = v
        else:
            m = torch.minimum(ramp(h, bounds[0], bounds[1], border[0])[:, None], ramp(w, bounds[2], bounds[3], border[1])[None, :])
            values[:, t0:t0 + frames, y0:y0 + h, x0:x0 + w] += v.to(BF).float() * m
            weight[t0:t0 + frames, y0:y0 + h, x0:x0 + w] += m

    def vae_blend_divide(values, weight):
        values /= weight

    # ---- decoder side
    def vae_latent_rows(z, mean, inv_std, grid, cp):
        C, T, H, W = z.shape
        g = grid.view(T, H + 2, W + 2, cp)
        v = (z.float() / inv_std.view(-1, 1, 1, 1)).to(BF).float() + mean.view(-1, 1, 1, 1)
        g[:, 1:-1, 1:-1, :C] = v.permute(1, 2, 3, 0).to(BF)

    def vae_upsample2x(src, dst, cp, frames_dst, h, w, halves=1):
        s = src[:(frames_dst // halves) * (h + 2) * (w + 2)].view(frames_dst // halves, h + 2, w + 2, halves, cp)[:, 1:-1, 1:-1]
        s = s.permute(0, 3, 1, 2, 4).reshape(frames_dst, h, w, cp)                # frame t' = half t' % halves of source frame t' // halves
        d = dst[:frames_dst * (2 * h + 2) * (2 * w + 2)].view(frames_dst, 2 * h + 2, 2 * w + 2, cp)
        d[:, 1:-1, 1:-1] = s.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)

    def vae_dup_up_add(x, main, cin, cout, factor_t, first_chunk, frames_out, h, w):
        t_in = (frames_out + (factor_t - 1 if first_chunk else 0)) // factor_t
        xv = x[:t_in * (h + 2) * (w + 2)].view(t_in, h + 2, w + 2, -1)[:, 1:-1, 1:-1, :cin].float().permute(3, 0, 1, 2)[None]
        add = o.dup_up3d(xv, cout, factor_t, 2, first_chunk)[0]
        m = main[:frames_out * (2 * h + 2) * (2 * w + 2)].view(frames_out, 2 * h + 2, 2 * w + 2, -1)
        m[:, 1:-1, 1:-1, :cout] = (m[:, 1:-1, 1:-1, :cout].float() + add.permute(1, 2, 3, 0)).to(BF)

    def vae_unpatchify(head, frames, h, w, values, weight, t0, y0, x0, bounds=(True,) * 4, border=(1, 1)):
        g = head[:frames * (h + 2) * (w + 2)].view(frames, h + 2, w + 2, -1)[:, 1:-1, 1:-1, :12].float().permute(3, 0, 1, 2)[None]
        v = o.unpatchify(g)[0]                                                    # [3, frames, 2h, 2w]
        vt = min(frames, values.shape[1] - t0)
        if weight is None:
            values[:, t0:t0 + vt, y0:y0 + 2 * h, x0:x0 + 2 * w] = v[:, :vt].clamp(-1, 1)
        else:
            m = torch.minimum(ramp(2 * h, bounds[0], bounds[1], border[0])[:, None], ramp(2 * w, bounds[2], bounds[3], border[1])[None, :])
            values[:, t0:t0 + vt, y0:y0 + 2 * h, x0:x0 + 2 * w] += v[:, :vt] * m
            weight[t0:t0 + vt, y0:y0 + 2 * h, x0:x0 + 2 * w] += m

    def vae_blend_finish(values, weight):
        values.copy_((values / weight).clamp(-1, 1))

    for name, fn in list(locals().items()):
        if callable(fn) and name not in ("ramp",) and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(ops, "context", lambda device: None)


def latents(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.fixture()
def encoder(monkeypatch):
    _emulated_ops(monkeypatch)
    from fairygen_b200 import vae, vae_encode
    cfg = vae.VAE38Config(z_dim=o.TINY.z_dim, dec_dim=o.TINY.dec_dim)
    enc = vae_encode.VAE38Encoder(cfg, "cpu", enc_dim=o.TINY.enc_dim)
    w = o.make_enc_weights(o.TINY, seed=0)
    enc.load_state_dict({"model." + k: v for k, v in w.items()})
    return enc, {k: v.to(BF).float() for k, v in w.items()}


def test_host_mirror_lists_the_reference_keys():
    from fairygen_b200 import vae, vae_encode
    assert vae_encode.enc_param_shapes(vae.VAE38) == o.enc_param_shapes(o.VAE38)
    assert vae_encode.enc_param_shapes(vae.VAE38Config(z_dim=8, dec_dim=16), 16) == o.enc_param_shapes(o.TINY)


@pytest.mark.parametrize("shape,seed", [((3, 1, 32, 48), 30), ((3, 9, 32, 32), 31), ((3, 5, 48, 32), 35)])
def test_encoder_orchestration_matches_the_oracle(encoder, shape, seed):
    """Image (one chunk), 9 frames (chunks of 1, 4, 4: both temporal down-samplings with their caches), 5 frames."""
    enc, w16 = encoder
    video = torch.tanh(latents(shape, seed)).to(BF)
    got = enc.encode([video])
    with torch.no_grad():
        want = o.encode(w16, o.TINY, [video.float()])
    assert got.shape == want.shape and got.dtype == BF
    print("encode", shape, rel(got.float(), want))
    assert rel(got.float(), want) < 2e-2, rel(got.float(), want)


def test_tiled_encode_matches_the_oracle_and_the_reference_golden(encoder):
    import os
    enc, w16 = encoder
    video = torch.tanh(latents((3, 1, 80, 96), 32)).to(BF)
    got = enc.encode([video], tiled=True, tile_size=(3, 4), tile_stride=(2, 2))
    with torch.no_grad():
        want = o.encode(w16, o.TINY, [video.float()], tiled=True, tile_size=(3, 4), tile_stride=(2, 2))
    assert rel(got.float(), want) < 2e-2
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "vae38.npz"))["encode_tiled"]
    assert rel(got.float(), torch.from_numpy(gold)) < 3e-2


@pytest.mark.parametrize("tiled", [False, True])
def test_decoder_orchestration_matches_the_oracle(monkeypatch, tiled):
    """The same emulated kernels under the decoder (whose real kernels are verified on the GPU): chunked decode of 3 latent
    frames and a tiled decode — the emulations state the kernels' contracts, the host logic around them is what is tested."""
    _emulated_ops(monkeypatch)
    from fairygen_b200 import vae
    dec = vae.VAE38Decoder(vae.VAE38Config(z_dim=o.TINY.z_dim, dec_dim=o.TINY.dec_dim), "cpu")
    w = o.make_weights(o.TINY, seed=0)
    dec.load_state_dict(w)
    w16 = {k: v.to(BF).float() for k, v in w.items()}
    z = (latents((1, 8, 2, 5, 5), 3) if tiled else latents((1, 8, 3, 3, 4), 1)).to(BF)
    kw = dict(tiled=True, tile_size=(3, 3), tile_stride=(2, 2)) if tiled else {}
    got = dec.decode(z, **kw)
    with torch.no_grad():
        want = o.decode(w16, o.TINY, z.float(), **kw)
    assert got.shape == want.shape
    assert rel(got.float(), want) < 2.5e-2, rel(got.float(), want)


def test_full_width_encoder_orchestration(monkeypatch):
    """The production widths (160 / 320 / 640 channels: 160 is NOT a multiple of the 64-channel grid padding, so every grid of
    the first stages carries zero padding channels; 96 output channels of the head; 48 latent channels) on one small image."""
    _emulated_ops(monkeypatch)
    from fairygen_b200 import vae, vae_encode
    enc = vae_encode.VAE38Encoder(vae.VAE38, "cpu")
    w = o.make_enc_weights(o.VAE38, seed=1)
    enc.load_state_dict(w)
    image = torch.tanh(latents((3, 1, 32, 32), 41)).to(BF)
    got = enc.encode([image])
    with torch.no_grad():
        want = o.encode({k: v.to(BF).float() for k, v in w.items()}, o.VAE38, [image.float()])
    assert got.shape == (1, 48, 1, 2, 2)
    assert rel(got.float(), want) < 2e-2, rel(got.float(), want)


def test_install_routes_pipe_vae_encode_and_decode(monkeypatch):
    """vae.install / vae_encode.install on a stand-in for the loaded reference VAE (its `model`, `z_dim`, `state_dict()`)."""
    _emulated_ops(monkeypatch)
    from fairygen_b200 import vae, vae_encode
    small = vae.VAE38Config(z_dim=48, dec_dim=16)
    monkeypatch.setattr(vae, "VAE38", small)
    monkeypatch.setattr(vae_encode, "VAE38", small)
    ocfg = o.VAE38Config(z_dim=48, dec_dim=16, enc_dim=16)
    sd = {"model." + k: v for k, v in {**o.make_weights(ocfg, seed=0), **o.make_enc_weights(ocfg, seed=0)}.items()}

    class Model:
        dim = 16

    class RefVAE(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model, self.z_dim = Model(), 48

        def state_dict(self, *a, **k):
            return dict(sd)

    class Pipe:
        device = "cpu"
        vae = RefVAE()

    pipe = Pipe()
    dec, enc = vae.install(pipe), vae_encode.install(pipe)
    assert pipe.vae.decode == dec.decode and pipe.vae.encode == enc.encode
    image = torch.tanh(latents((3, 1, 32, 32), 40)).to(BF)
    z = pipe.vae.encode([image], device="cpu")                                   # PIPE:495
    assert z.shape == (1, 48, 1, 2, 2)
    video = pipe.vae.decode(z, device="cpu", tiled=False)                        # PIPE:323
    assert video.shape == (1, 3, 1, 32, 32) and float(video.float().abs().max()) <= 1
    with torch.no_grad():
        w16 = {k[len("model."):]: v.to(BF).float() for k, v in sd.items()}
        assert rel(z.float(), o.encode(w16, ocfg, [image.float()])) < 2e-2
    pipe.vae = None
    with pytest.raises(ValueError):
        vae_encode.install(pipe)


def test_encoder_rejects_bad_input(encoder):
    enc, _ = encoder
    with pytest.raises(ValueError):
        enc.encode([torch.zeros(3, 2, 32, 32)])          # 2 frames: not 1 + 4k
    with pytest.raises(ValueError):
        enc.encode([torch.zeros(3, 1, 24, 32)])          # height not a multiple of 16
    with pytest.raises(ValueError):
        enc.encode([torch.zeros(4, 1, 32, 32)])


def test_window_assignment_covers_every_window_once_and_balances():
    """assign_windows (the multi-GPU split of the tiled decode, tiled_decode of wan_video_vae.py:1081-1152 made parallel): every
    window goes to exactly one rank, and no rank carries more than the mean load plus one window."""
    from fairygen_b200 import vae
    for (H, W, size, stride) in ((44, 80, (34, 34), (18, 16)), (30, 52, (34, 34), (18, 16)), (6, 7, (3, 4), (2, 3))):
        tasks = vae.tile_tasks(H, W, size, stride)
        area = lambda t: (min(t[1], H) - t[0]) * (min(t[3], W) - t[2])  # noqa: E731
        for world in (1, 2, 3, 4, 8):
            mine = vae.assign_windows(tasks, H, W, world)
            assert len(mine) == world and sorted(t for m in mine for t in m) == sorted(tasks)
            loads = [sum(area(t) for t in m) for m in mine]
            assert max(loads) <= sum(loads) / world + max(area(t) for t in tasks), (H, W, world, loads)


def _windows_worker(rank, world, port, out_dir):
    import os
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo)
    sys.path.insert(0, os.path.join(repo, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from fairygen_b200 import vae
    from test_vae_host import _emulated_ops as emu

    class _Patch:
        @staticmethod
        def setattr(obj, name, value):
            setattr(obj, name, value)

    emu(_Patch)
    dec = vae.VAE38Decoder(vae.VAE38Config(z_dim=o.TINY.z_dim, dec_dim=o.TINY.dec_dim), "cpu")
    dec.load_state_dict(o.make_weights(o.TINY, seed=0))
    z = torch.randn((1, o.TINY.z_dim, 2, 6, 7), generator=torch.Generator().manual_seed(3)).to(BF)
    kw = dict(tiled=True, tile_size=(3, 4), tile_stride=(2, 3))
    alone = dec.decode(z, **kw)
    launches_alone, dec.kernel_launches = dec.kernel_launches, 0
    shared = dec.decode(z, group=dist.group.WORLD, **kw)
    torch.save({"err": float((shared.float() - alone.float()).abs().max()), "launches": (dec.kernel_launches, launches_alone)},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_decode_windows_over_two_ranks(tmp_path):
    """CPU twin (gloo) of tests/test_sp_gpu.py::test_vae_windows_over_two_gpus: the windows of a tiled decode spread over two
    ranks, one sum all-reduce of the blended video and its weights — every rank ends with the single-rank result, having
    decoded only its share of the windows."""
    import os
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_windows_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    shares = []
    for r in range(2):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["err"] < 1e-2, res                 # bf16 output; fp32 sums of 2-4 contributions in another order
        shares.append(res["launches"][0])
        assert res["launches"][0] < res["launches"][1], res
    assert abs(sum(shares) - res["launches"][1]) <= 2, (shares, res)     # together: the windows once (+ one blend_finish per rank)
