"""CPU twin of tests/test_reference_pipeline_gpu.py: the UNMODIFIED reference pipeline (baseline/_ref) drives the drop-in boundary.
``WanVideoPipeline.__call__`` (wan_video.py:172-329) — units, scheduler, denoise loop, the ~40-key ``model_fn`` call — runs once
with the reference's own ``model_fn_wan_video`` and once after ``fairygen_b200.install(pipe)``, here on the CPU: the product's
host code (install, model_fn_wan_video, engine_for / re-pack, WanDiTEngine.forward, FlowMatchScheduler) runs for real, every
kernel is the plain-torch statement of its contract from tests/test_engine_host.py.  What this pins without a GPU: the keyword
surface, the swallowed keys, per-token timesteps, the scheduler swap, ``load_lora`` after ``install``."""
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
BF = torch.bfloat16


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.fixture()
def env(monkeypatch):
    from baseline import ref_loader as rl

    if not rl.available():
        pytest.skip("baseline/_ref not installed (python baseline/install_ref.py in the build container)")
    import fairygen_b200 as fg
    from fairygen_b200 import model_fn as mf
    from fairygen_b200 import ops, scheduler
    from oracle import wan_dit_oracle as o
    from test_engine_host import _emulated_ops

    rl.load()
    rl.select_attention_backend("cpu")
    _emulated_ops(monkeypatch)

    class HostEngine(fg.WanDiTEngine):                       # the product's constructor refuses the CPU; same fields by hand
        def __init__(self, cfg, device="cpu", sp=None):
            self.cfg, self.device, self.ctx, self.sp = cfg, torch.device("cpu"), None, sp
            self.rope_tab = torch.from_numpy(ops.rope_table(cfg.head_dim))
            self._init_state()

    monkeypatch.setattr(mf, "WanDiTEngine", HostEngine)
    monkeypatch.setattr(mf, "_require_device_weights", lambda device: None)

    def step_fused(self, latents, noise_pos, noise_neg, cfg_scale, index, first_frame_latents=None, to_final=False):
        ops.cfg_fm_step(latents, noise_pos, noise_neg, first_frame_latents, float(cfg_scale), self.sigma_delta(index, to_final))
        return latents

    monkeypatch.setattr(scheduler.FlowMatchScheduler, "step_fused", step_fused)      # without the CUDA-only guard
    return fg, mf, o, rl


def _pipes(o, rl, seed=0):
    w = {k: v.to(BF) for k, v in o.make_weights(o.TINY, seed=seed).items()}
    pipes = []
    for _ in range(2):
        dit = rl.build_wan_model(o.TINY, state_dict={k: v.clone() for k, v in w.items()})
        pipes.append(rl.build_pipeline(dit, "cpu", BF, text_len=32))
    return pipes


def _call(pipe, **kw):
    from PIL import Image

    img = Image.fromarray((np.random.RandomState(0).rand(64, 64, 3) * 255).astype(np.uint8))
    args = dict(prompt="a paper boat drifts down the gutter after the rain", negative_prompt="blurry, static", input_image=img, seed=3,
                height=64, width=64, num_frames=9, num_inference_steps=4, tiled=False, output_type="floatpoint",
                progress_bar_cmd=lambda x: x)
    args.update(kw)
    video = pipe(**args)
    return video, pipe.vae.last_latents


def test_install_drives_the_unmodified_pipeline_on_the_host(env):
    fg, mf, o, rl = env
    ref_pipe, our_pipe = _pipes(o, rl)
    calls = []
    ref_fn = ref_pipe.model_fn

    def recording_model_fn(**kwargs):        # wraps, does not modify, the reference's model_fn
        out = ref_fn(**kwargs)
        calls.append(({k: (v.clone() if torch.is_tensor(v) else v) for k, v in kwargs.items()}, out.clone()))
        return out

    ref_pipe.model_fn = recording_model_fn
    video_ref, lat_ref = _call(ref_pipe)
    assert len(calls) == 8 and len(calls[0][0]) > 40            # 4 steps x (positive, negative); the full keyword soup

    fg.install(our_pipe)
    assert our_pipe.model_fn is fg.model_fn_wan_video and type(our_pipe.scheduler).__module__ == "fairygen_b200.scheduler"
    video, lat = _call(our_pipe)
    assert lat.shape == lat_ref.shape and lat.dtype == lat_ref.dtype
    print(f"4-step pipeline call on the host: latents rel L2 {rel_l2(lat, lat_ref):.3e}")
    assert rel_l2(lat, lat_ref) < 3e-2                          # north star: <= 3e-2 after a schedule
    assert torch.equal(lat[:, :, 0:1], lat_ref[:, :, 0:1])      # first frame restored from the (shared) stub VAE encode
    assert video.shape == video_ref.shape

    # every recorded reference call, replayed through the drop-in with the reference's exact kwargs: <= 1e-2 per forward
    eng = mf.engine_for(our_pipe.dit)
    misses = eng.ctx_cache_misses
    assert misses == 2                                          # the two prompts were projected once each for the whole schedule
    worst = 0.0
    for kwargs, want in calls:
        kwargs = dict(kwargs)
        kwargs["dit"] = our_pipe.dit
        got = fg.model_fn_wan_video(**kwargs)
        assert got.shape == want.shape and got.dtype == want.dtype
        worst = max(worst, rel_l2(got, want))
    print(f"per-call worst rel L2 {worst:.3e}")
    assert worst < 1e-2, worst
    assert eng.ctx_cache_misses == misses and eng.ctx_cache_content_hits >= 2   # the replayed (cloned) contexts hit by content


def test_load_lora_after_install_repacks_on_the_host(env):
    fg, mf, o, rl = env
    from fairygen_b200 import synthetic

    ref_pipe, our_pipe = _pipes(o, rl, seed=1)
    fg.install(our_pipe)
    eng = mf.engine_for(our_pipe.dit)
    cfg = fg.WanDiTConfig.from_module(our_pipe.dit)
    lat, z0, cp, cn = o.make_inputs(o.TINY, (1, 48, 3, 8, 8), text_len=32, live_text=8)
    lat, cp, ts = lat.to(BF), cp.to(BF), torch.tensor([900.0], dtype=BF)
    kw = dict(latents=lat, timestep=ts, context=cp, fuse_vae_embedding_in_latents=True)
    with torch.no_grad():
        before = our_pipe.model_fn(dit=our_pipe.dit, **kw)
        lora = synthetic.random_lora(cfg, rank=8, seed=5, device="cpu")
        for pipe in (ref_pipe, our_pipe):
            pipe.load_lora(pipe.dit, state_dict={k: v.clone() for k, v in lora.items()}, alpha=1.0)   # LORA:44-62, in place
        want = ref_pipe.model_fn(dit=ref_pipe.dit, **kw)
        got = our_pipe.model_fn(dit=our_pipe.dit, **kw)
        assert mf.engine_for(our_pipe.dit) is eng                # same engine, re-packed
        assert rel_l2(got, want) < 1e-2 and rel_l2(got, before) > 1e-3
        # the keys the hot path does not serve are refused, not ignored (no silent fallback to the reference kernels)
        with pytest.raises(NotImplementedError):
            our_pipe.model_fn(dit=our_pipe.dit, vace_context=torch.zeros(1), **kw)
        # merged CFG (PIPE:785-803): a batch of two contexts is two forwards
        both = our_pipe.model_fn(dit=our_pipe.dit, latents=lat, timestep=ts, context=torch.cat([cp, cn.to(BF)]),
                                 fuse_vae_embedding_in_latents=True, cfg_merge=True)
        assert both.shape[0] == 2 and torch.equal(both[0:1], got)
