"""The oracle (oracle/wan_dit_oracle.py) against the golden vectors produced by the REAL reference
(oracle/make_golden.py, committed under tests/golden/).  CPU only."""
import numpy as np
import torch

from oracle import wan_dit_oracle as o

T = torch.from_numpy


def close(a, b, tol=1e-5):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), f"max abs err {err}"


def test_tiny_forward_matches_reference(golden):
    g = golden("tiny_forward")
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    lat, z0, cp, cn = o.make_inputs(cfg, (1, cfg.in_dim, 3, 8, 8), text_len=32, live_text=8)
    with torch.no_grad():
        close(o.dit_forward(w, cfg, lat, torch.tensor([900.0]), cp, True), g["fused_t900"])
        close(o.dit_forward(w, cfg, lat, torch.tensor([37.0]), cp, True), g["fused_t37"])
        close(o.dit_forward(w, cfg, lat, torch.tensor([900.0]), cn, False), g["plain_t900"])
        lat2, _, cp2, _ = o.make_inputs(cfg, (1, cfg.in_dim, 2, 6, 10), text_len=24, live_text=24)
        close(o.dit_forward(w, cfg, lat2, torch.tensor([500.0]), cp2, True), g["ragged_fused_t500"])


def test_tiny_forward_bf16_close_to_reference_bf16(golden):
    g = golden("tiny_forward")
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0, dtype=torch.bfloat16)
    lat, z0, cp, cn = o.make_inputs(cfg, (1, cfg.in_dim, 3, 8, 8), dtype=torch.bfloat16, text_len=32, live_text=8)
    with torch.no_grad():
        out = o.dit_forward(w, cfg, lat, torch.tensor([900.0]).bfloat16(), cp, True).float()
    ref = T(g["fused_t900_bf16"])
    assert (out - ref).norm() / ref.norm() < 1e-6  # same ops, same dtype -> same result on the same CPU kernels


def test_ops_match_reference(golden):
    g = golden("ops")
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    x = T(g["x"])
    close(o.sinusoidal_embedding_1d(256, torch.tensor([0.0, 1.0, 37.0, 500.0, 996.0, 1000.0])), g["sinusoid"], 1e-6)
    close(o.sinusoidal_embedding_1d(256, torch.tensor([0.0, 996.0, 1000.0]).bfloat16()).float(), g["sinusoid_bf16"], 0)
    freqs = o.rope_freqs(o.rope_tables_3d(cfg.head_dim), 2, 3, 5)
    close(freqs.real, g["freqs_real"], 1e-12)
    close(freqs.imag, g["freqs_imag"], 1e-12)
    close(o.rope_apply(x, freqs, cfg.num_heads), g["rope"], 1e-6)
    close(o.rms_norm(x, w["blocks.1.self_attn.norm_q.weight"], cfg.eps), g["rmsnorm"], 1e-6)
    close(o.modulate(x, x.flip(1) * 0.1, x.flip(2) * 0.2), g["modulate"], 1e-6)
    ctx = T(g["ctx"])
    with torch.no_grad():
        close(o.self_attention(w, "blocks.1.self_attn.", x, freqs, cfg), g["self_attn"])
        close(o.cross_attention(w, "blocks.1.cross_attn.", x, ctx, cfg), g["cross_attn"])
        close(o.dit_block(w, 1, x, ctx, T(g["t_mod_tok"]), freqs, cfg), g["block_tok"])
        close(o.dit_block(w, 1, x, ctx, T(g["t_mod_one"]), freqs, cfg), g["block_one"])
        close(o.head(w, x, T(g["t_tok"]), cfg), g["head_tok"])
        close(o.head(w, x, T(g["t_one"]), cfg), g["head_one"])
        close(o.unpatchify(o.head(w, x, T(g["t_tok"]), cfg), (2, 3, 5), cfg), g["unpatchify"])
        tokens, grid = o.patch_embed(w, T(g["lat"]), cfg)
        ref = T(g["patchify"])  # (1, D, f, h, w)
        assert grid == tuple(ref.shape[2:])
        close(tokens, ref.flatten(2).transpose(1, 2))


def test_scheduler_matches_reference(golden):
    g = golden("scheduler")
    for n, shift in ((50, 5.0), (8, 3.0)):
        sig, ts = o.flow_match_schedule(n, 1.0, shift)
        assert np.array_equal(sig.numpy(), g[f"sigmas_{n}"])
        assert np.array_equal(ts.numpy(), g[f"timesteps_{n}"])
        assert np.array_equal(ts.to(torch.bfloat16).float().numpy(), g[f"timesteps_bf16_{n}"])
    sig, _ = o.flow_match_schedule(50, 1.0, 5.0)
    sample, npos, nneg, z0 = (T(g[k]).bfloat16() for k in ("sample", "npos", "nneg", "z0"))
    for i in (0, 17, 49):
        nxt = o.flow_match_step(o.cfg_combine(npos, nneg, 5.0), i, sample, sig)
        assert nxt.dtype == torch.bfloat16
        nxt[:, :, 0:1] = z0
        assert torch.equal(nxt.float(), T(g[f"step_{i}"]))


def test_lora_fuse_matches_reference(golden):
    g = golden("lora")
    cfg = o.TINY
    lora = o.make_lora(cfg, rank=8, seed=2)
    assert len(o.lora_target_names(lora)) == 20
    for dtype, tag, tol in ((torch.float32, "f32", 1e-6), (torch.bfloat16, "bf16", 0)):
        fused = o.fuse_lora(o.make_weights(cfg, seed=0, dtype=dtype), lora, 1.0)
        for key in g:
            if not key.startswith(tag + ":") or key.endswith(":sum"):
                continue
            name = key.split(":", 1)[1]
            close(fused[name].float()[:32], g[key], tol)
            assert abs(float(fused[name].double().sum()) - float(g[key + ":sum"])) <= 1e-3 * max(1.0, abs(float(g[key + ":sum"])))


def test_usp_glue_matches_reference(golden):
    g = golden("usp")
    cfg = o.TINY
    x = T(g["x"])
    freqs = o.rope_freqs(o.rope_tables_3d(cfg.head_dim), 2, 3, 5)
    chunks, pad = o.sp_chunk_pad(x, 4)
    assert pad == 2
    for r in range(4):
        close(chunks[r], g[f"chunk_rank{r}"], 0)
        close(o.sp_rope_apply(chunks[r], freqs, cfg.num_heads, r, 4), g[f"rope_rank{r}"], 1e-6)


def test_ulysses_emulation_equals_single_rank():
    """Scatter-heads/gather-sequence attention on virtual ranks == plain attention when padded keys are masked."""
    g = torch.Generator().manual_seed(3)
    heads, s, world = 4, 30, 4
    q, k, v = (torch.randn(1, s, heads * 16, generator=g) for _ in range(3))
    full = o.attention(q, k, v, heads)
    qs, pad = o.sp_chunk_pad(q, world)
    ks, _ = o.sp_chunk_pad(k, world)
    vs, _ = o.sp_chunk_pad(v, world)
    outs = o.ulysses_attention(qs, ks, vs, heads, valid_tokens=s)
    got = torch.cat(outs, dim=1)[:, :s]
    assert (got - full).abs().max() < 1e-5
    # the reference does NOT mask the zero-padded rows: result differs when S % P != 0
    ref_like = torch.cat(o.ulysses_attention(qs, ks, vs, heads, valid_tokens=None), dim=1)[:, :s]
    assert (ref_like - full).abs().max() > 1e-4


def test_denoise_loop_matches_reference(golden):
    g = golden("denoise")
    cfg = o.TINY
    w = o.make_weights(cfg, seed=0)
    lat, z0, cp, cn = o.make_inputs(cfg, (1, cfg.in_dim, 3, 8, 8), text_len=32, live_text=8)
    lat[:, :, 0:1] = z0
    with torch.no_grad():
        out = o.denoise(w, cfg, lat, cp, cn, z0, num_inference_steps=4, cfg_scale=5.0, shift=5.0)
    close(out, g["final"], 1e-5)


def test_counted_flops_match_survey():
    assert abs(o.counted_flops(o.TI2V_5B, 27280) / 5.1641e14 - 1) < 1e-3
    assert abs(o.counted_flops(o.TI2V_5B, 320) / 2.8772e12 - 1) < 1e-3
