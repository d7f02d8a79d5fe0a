"""End-to-end parity of the CUDA path on the B200: engine / drop-in model_fn / denoise loop against
(a) golden outputs of the REAL reference (tests/golden, fp32 and bf16 CPU runs) and (b) the oracle's
restatement executed on the GPU in bf16 (= the reference's bf16 op chain on cuBLAS/SDPA) and fp32.
Tolerances are the north star's: latent relative L2 <= 1e-2 per forward/step, <= 3e-2 after a schedule."""
import types

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
PER_STEP_TOL = 1e-2
SCHEDULE_TOL = 3e-2


@pytest.fixture(scope="module")
def env():
    import fairygen_b200
    from fairygen_b200 import ops
    from oracle import wan_dit_oracle as o
    torch.cuda.set_device(0)
    ops.context(torch.device("cuda", 0))
    return fairygen_b200, o


def tiny_engine(fg, o, weights=None):
    cfg = fg.WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    eng = fg.WanDiTEngine(cfg, "cuda")
    eng.load_state_dict(weights if weights is not None else o.make_weights(o.TINY, seed=0))
    return eng


def test_tiny_forward_vs_reference_golden(env, golden):
    fg, o = env
    g = golden("tiny_forward")
    eng = tiny_engine(fg, o)
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 8, 8), text_len=32, live_text=8)
    cases = [("fused_t900", lat, 900.0, cp, True), ("fused_t37", lat, 37.0, cp, True), ("plain_t900", lat, 900.0, cn, False)]
    lat2, _, cp2, _ = o.make_inputs(o.TINY, (1, 48, 2, 6, 10), text_len=24, live_text=24)
    cases.append(("ragged_fused_t500", lat2, 500.0, cp2, True))
    for name, la, t, ctx, fuse in cases:
        out = eng.forward(la.cuda(), torch.tensor([t]), ctx.cuda(), fuse)
        fg.ops.sync_check()
        assert out.shape == la.shape and out.dtype == la.dtype
        err = rel_l2(out, torch.from_numpy(g[name]))
        assert err < PER_STEP_TOL, f"{name}: rel L2 {err:.3e} vs the reference fp32 output"
    out = eng.forward(lat.cuda().to(BF), torch.tensor([900.0]), cp.cuda().to(BF), True)
    assert out.dtype == BF
    assert rel_l2(out, torch.from_numpy(g["fused_t900_bf16"])) < PER_STEP_TOL   # reference bf16 path (CPU kernels)


def test_drop_in_model_fn_and_lora_repack(env, golden):
    fg, o = env
    from helpers import WanContainer
    g = golden("tiny_forward")
    w = o.make_weights(o.TINY, seed=0)
    dit = WanContainer(o.TINY)
    dit.load_state_dict(w, strict=True)
    dit = dit.to("cuda", BF)
    pipe = types.SimpleNamespace(dit=dit, model_fn=None, scheduler=None)
    fg.install(pipe)
    assert pipe.model_fn is fg.model_fn_wan_video and type(pipe.scheduler).__name__ == "FlowMatchScheduler"
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 8, 8), dtype=BF, text_len=32, live_text=8)
    # called exactly like wan_video.py:296 — models, shared inputs (incl. unrelated keys), posi inputs, timestep
    models = dict(dit=dit, motion_controller=None, vace=None, animate_adapter=None, vap=None)
    shared = dict(latents=lat.cuda(), input_image=None, seed=1, height=128, width=128, num_frames=9, cfg_scale=5.0,
                  cfg_merge=False, tiled=True, tile_size=(30, 52), sigma_shift=5.0, fuse_vae_embedding_in_latents=True,
                  first_frame_latents=z0.cuda(), rand_device="cpu", use_unified_sequence_parallel=False)
    out = pipe.model_fn(**models, **shared, context=cp.cuda(), timestep=torch.tensor([900.0], dtype=BF, device="cuda"))
    assert out.shape == lat.shape and out.dtype == BF and out.is_cuda
    assert rel_l2(out, torch.from_numpy(g["fused_t900"])) < PER_STEP_TOL
    # pipe.load_lora fuses into the container in place (general.py:44-62); the engine must notice and re-pack
    lora = o.make_lora(o.TINY, rank=8, seed=2)
    fused = o.fuse_lora(w, lora, 1.0)
    dit.load_state_dict({k: v.to(BF) for k, v in fused.items()}, strict=True)
    out2 = pipe.model_fn(**models, **shared, context=cp.cuda(), timestep=torch.tensor([900.0], dtype=BF, device="cuda"))
    with torch.no_grad():
        want = o.dit_forward(fused, o.TINY, lat.float(), torch.tensor([900.0]), cp.float(), True)
    assert rel_l2(out2, want) < PER_STEP_TOL
    assert rel_l2(out2, out) > 1e-3   # the LoRA really changed the result
    # merged CFG: a batch of two contexts -> two predictions (wan_video.py:1240-1243)
    both = pipe.model_fn(**models, **shared, context=torch.cat([cp, cn]).cuda(), timestep=torch.tensor([900.0], dtype=BF, device="cuda"))
    assert both.shape[0] == 2 and torch.equal(both[0:1], out2)


def test_denoise_loop_vs_reference_golden(env, golden):
    fg, o = env
    g = golden("denoise")
    eng = tiny_engine(fg, o)
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 8, 8), text_len=32, live_text=8)
    den = fg.WanDenoiser(eng, num_inference_steps=4, cfg_scale=5.0, sigma_shift=5.0)
    out = den(lat, cp, cn, z0)
    fg.ops.sync_check()
    err = rel_l2(out, torch.from_numpy(g["final"]))
    assert err < SCHEDULE_TOL, f"4-step CFG denoise: rel L2 {err:.3e}"
    assert torch.equal(out[:, :, 0:1].cpu(), z0.to(BF))   # first frame restored exactly (wan_video.py:308-309)


def _real_dim_case(fg, o, layers, shape, text_len=512, seed=0):
    from fairygen_b200 import synthetic
    cfg = fg.WanDiTConfig(num_layers=layers)
    ocfg = o.DiTConfig(num_layers=layers)
    sd = synthetic.random_state_dict(cfg, seed=seed, device="cuda", dtype=BF, lora_rank=32)
    eng = fg.WanDiTEngine(cfg, "cuda")
    eng.load_state_dict(sd)
    lat, z0, cp, cn = synthetic.synthetic_inputs(cfg, shape, text_len=text_len, pin=False)
    return cfg, ocfg, sd, eng, lat.cuda(), cp.cuda()


def test_real_dims_two_layers_vs_oracle_bf16_and_fp32(env):
    """D=3072, F=14336, 24 heads, merged rank-32 LoRA, BASELINE config-1 shape (S=320, 512 text tokens)."""
    fg, o = env
    cfg, ocfg, sd, eng, lat, cp = _real_dim_case(fg, o, 2, (1, 48, 5, 16, 16))
    ts = torch.tensor([996.0], device="cuda", dtype=BF)
    out = eng.forward(lat, ts, cp, True)
    with torch.no_grad():
        ref16 = o.dit_forward(sd, ocfg, lat, ts, cp, True)
        ref32 = o.dit_forward({k: v.float() for k, v in sd.items()}, ocfg, lat.float(), ts.float(), cp.float(), True)
    e_mine, e_ref, e_pair = rel_l2(out, ref32), rel_l2(ref16, ref32), rel_l2(out, ref16)
    print(f"2 layers S=320: ours-vs-fp32 {e_mine:.3e}  ref_bf16-vs-fp32 {e_ref:.3e}  ours-vs-ref_bf16 {e_pair:.3e}")
    assert e_pair < PER_STEP_TOL and e_mine < PER_STEP_TOL
    assert e_mine < 2.0 * e_ref + 1e-3   # our error is of the size of the reference's own bf16 error


def test_full_model_config1_vs_oracle_bf16(env):
    """All 30 layers at BASELINE config 1 (17 frames 256x256 -> S=320): ours vs the reference bf16 op chain."""
    fg, o = env
    cfg, ocfg, sd, eng, lat, cp = _real_dim_case(fg, o, 30, (1, 48, 5, 16, 16))
    ts = torch.tensor([996.0], device="cuda", dtype=BF)
    out = eng.forward(lat, ts, cp, True)
    out_again = eng.forward(lat, ts, cp, True)
    assert torch.equal(out, out_again), "forward must be deterministic"
    with torch.no_grad():
        ref16 = o.dit_forward(sd, ocfg, lat, ts, cp, True)
    err = rel_l2(out, ref16)
    print(f"30 layers S=320: ours-vs-ref_bf16 {err:.3e}")
    assert torch.isfinite(out.float()).all() and err < PER_STEP_TOL


def test_headline_shape_full_model_vs_oracle_bf16(env):
    """704x1280x121 -> latent (1,48,31,44,80), S = 27 280, 30 layers, CFG-style pair of contexts: parity with the
    reference bf16 op chain executed by torch on the same GPU, plus one fused scheduler step."""
    fg, o = env
    cfg, ocfg, sd, eng, lat, cp = _real_dim_case(fg, o, 30, (1, 48, 31, 44, 80))
    ts = torch.tensor([996.0], device="cuda", dtype=BF)
    out = eng.forward(lat, ts, cp, True)
    fg.ops.sync_check()
    assert torch.isfinite(out.float()).all()
    with torch.no_grad():
        ref16 = o.dit_forward(sd, ocfg, lat, ts, cp, True)
    err = rel_l2(out, ref16)
    print(f"30 layers S=27280: ours-vs-ref_bf16 {err:.3e}")
    assert err < PER_STEP_TOL


def test_on_gpu_lora_fuse_matches_reference_fuse(env):
    """lora_io.fuse_into_engine (one fgb_lora_merge per adapted Linear on the packed weights) vs the reference's
    GeneralLoRALoader.fuse_lora_to_base_model as restated by the oracle (pinned by tests/golden/lora.npz)."""
    fg, o = env
    from fairygen_b200 import lora_io
    w = o.make_weights(o.TINY, seed=0)
    lora = o.make_lora(o.TINY, rank=32, seed=2)
    ref_eng = tiny_engine(fg, o, o.fuse_lora({k: v.to(BF) for k, v in w.items()}, {k: v.to(BF) for k, v in lora.items()}, alpha=1.0))
    eng = tiny_engine(fg, o, w)
    assert lora_io.fuse_into_engine(eng, lora, alpha=1.0) == 20
    fg.ops.sync_check()
    for b, rb in zip(eng.blocks, ref_eng.blocks):
        for name in ("wqkv", "wo", "cwq", "cwkv", "cwo", "w1", "w2"):
            assert rel_l2(getattr(b, name), getattr(rb, name)) < 3e-3, name   # fp32-accumulated vs bf16 mm + bf16 add
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 8, 8), text_len=32, live_text=8)
    ts = torch.tensor([900.0])
    assert rel_l2(eng.forward(lat.cuda(), ts, cp.cuda(), True), ref_eng.forward(lat.cuda(), ts, cp.cuda(), True)) < 5e-3


@pytest.mark.parametrize("real_dims", [False, True])
def test_full_50_step_schedule_vs_reference_bf16_chain(env, real_dims):
    """North star: latent relative L2 <= 3e-2 after the FULL 50-step CFG schedule (wan_video.py:282-309) against the reference's
    bf16 path — the oracle's restatement run in bf16 on the same GPU (cuBLAS / SDPA) — and both against the fp32 oracle, to show
    that our accumulated error is of the size of the reference's own bf16 error. Tiny model (S = 48) and TI2V-5B block shapes
    (D = 3072, 24 heads, F = 14336, merged rank-32 LoRA; 2 layers, S = 320)."""
    fg, o = env
    if real_dims:
        cfg, ocfg, sd, eng, lat, cp = _real_dim_case(fg, o, 2, (1, 48, 5, 16, 16))
        from fairygen_b200 import synthetic
        _, z0, _, cn = synthetic.synthetic_inputs(cfg, (1, 48, 5, 16, 16), text_len=512, pin=False)
        z0, cn = z0.cuda(), cn.cuda()
        w = sd
    else:
        ocfg = o.TINY
        w = {k: v.cuda() for k, v in o.make_weights(ocfg, seed=0).items()}
        eng = tiny_engine(fg, o)
        lat, z0, cp, cn = (t.cuda() for t in o.make_inputs(ocfg, (1, 48, 3, 8, 8), text_len=32, live_text=8))
    den = fg.WanDenoiser(eng, num_inference_steps=50, cfg_scale=5.0, sigma_shift=5.0)
    ours = den(lat, cp, cn, z0)
    fg.ops.sync_check()
    with torch.no_grad():
        w16 = {k: v.to(BF) for k, v in w.items()}
        start16 = lat.to(BF).clone()
        start16[:, :, 0:1] = z0.to(BF)
        ref16 = o.denoise(w16, ocfg, start16, cp.to(BF), cn.to(BF), z0.to(BF), 50, 5.0, 5.0)
        w32 = {k: v.to(BF).float() for k, v in w.items()}
        start32 = lat.to(BF).float().clone()
        start32[:, :, 0:1] = z0.to(BF).float()
        ref32 = o.denoise(w32, ocfg, start32, cp.to(BF).float(), cn.to(BF).float(), z0.to(BF).float(), 50, 5.0, 5.0)
    e_ours16, e_ours32, e_ref = rel_l2(ours, ref16), rel_l2(ours, ref32), rel_l2(ref16, ref32)
    print(f"50-step schedule ({'real dims' if real_dims else 'tiny'}): ours-vs-ref_bf16 {e_ours16:.3e}  ours-vs-fp32 {e_ours32:.3e}  "
          f"ref_bf16-vs-fp32 {e_ref:.3e}")
    assert torch.isfinite(ours.float()).all()
    assert e_ours16 < SCHEDULE_TOL
    assert e_ours32 < max(SCHEDULE_TOL, 2.0 * e_ref)
