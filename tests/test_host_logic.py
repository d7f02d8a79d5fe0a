"""Host-side logic of the product that needs no GPU: schedule, RoPE table, config inference, partitioning,
the model_fn boundary's argument handling.  Golden vectors come from the real reference."""
import numpy as np
import pytest
import torch

from fairygen_b200 import TI2V_5B, WanDiTConfig, counted_flops, ops, sp, synthetic
from fairygen_b200.scheduler import FlowMatchScheduler


def test_schedule_bit_exact_vs_reference(golden):
    g = golden("scheduler")
    s = FlowMatchScheduler("Wan")
    for n, shift in ((50, 5.0), (8, 3.0)):
        s.set_timesteps(n, denoising_strength=1.0, shift=shift)
        assert np.array_equal(s.sigmas.numpy(), g[f"sigmas_{n}"])
        assert np.array_equal(s.timesteps.numpy(), g[f"timesteps_{n}"])
    s.set_timesteps(50, shift=5.0)
    assert s.sigma_delta(49) == -float(s.sigmas[49])          # last step goes to sigma = 0 (FM:149-150)
    assert s.sigma_delta(3) == float(s.sigmas[4] - s.sigmas[3])
    assert s._index(s.timesteps[17]) == 17 and s._index(torch.tensor(1000.0)) == 0


def test_only_wan_template():
    with pytest.raises(NotImplementedError):
        FlowMatchScheduler("FLUX.1")


def test_scheduler_step_has_no_cpu_fallback():
    s = FlowMatchScheduler("Wan")
    s.set_timesteps(4)
    x = torch.zeros(1, 48, 2, 4, 4, dtype=torch.bfloat16)
    with pytest.raises(NotImplementedError):
        s.step(x, s.timesteps[0], x)


def test_rope_table_matches_reference_phasors(golden):
    g = golden("ops")
    tab = ops.rope_table(128)
    assert tab.shape == (1024, 64, 2) and tab.dtype == np.float32
    f, h, w = 2, 3, 5
    real, imag = g["freqs_real"][:, 0], g["freqs_imag"][:, 0]  # [f*h*w, 64]
    for t in range(f * h * w):
        fi, hi, wi = t // (h * w), (t // w) % h, t % w
        pos = np.array([fi] * 22 + [hi] * 21 + [wi] * 21)
        assert np.allclose(tab[pos, np.arange(64), 0], real[t], atol=1e-7)
        assert np.allclose(tab[pos, np.arange(64), 1], imag[t], atol=1e-7)


def test_config_inference_and_flops():
    tiny = WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=2)
    shapes = synthetic.param_shapes(tiny)
    sd = {k: torch.empty(v, device="meta") for k, v in shapes.items()}
    assert WanDiTConfig.from_state_dict(sd) == tiny
    assert len(synthetic.param_shapes(TI2V_5B)) == 15 + 30 * 27
    n_params = sum(int(np.prod(s)) for s in synthetic.param_shapes(TI2V_5B).values())
    assert n_params == 4_999_787_712  # SURVEY §8 [probed]
    assert len(list(synthetic.lora_targets(TI2V_5B))) == 300
    assert abs(counted_flops(TI2V_5B, 27280) / 5.1641e14 - 1) < 1e-3
    with pytest.raises(ValueError):
        WanDiTConfig(dim=512, num_heads=8).validate()  # head_dim 64


def test_latent_shapes_of_the_baseline_configs():
    assert synthetic.latent_shape(TI2V_5B, 704, 1280, 121) == (1, 48, 31, 44, 80)   # S = 31*22*40 = 27 280
    assert synthetic.latent_shape(TI2V_5B, 256, 256, 17) == (1, 48, 5, 16, 16)      # S = 320
    assert synthetic.latent_shape(TI2V_5B, 480, 832, 81) == (1, 48, 21, 30, 52)     # S = 8190


def test_sequence_partition():
    # headline shape divides exactly for 1/2/4/8 ranks
    for world in (1, 2, 4, 8):
        rows = [sp.partition(27280, world, r) for r in range(world)]
        assert sum(r[2] for r in rows) == 27280 and all(r[0] * world == 27280 for r in rows)
    # S = 8190 needs 2 pad rows at 4 and 8 ranks
    assert sp.partition(8190, 4, 3) == (2048, 6144, 2046)
    assert sp.partition(8190, 8, 7) == (1024, 7168, 1022)
    # first-frame rows: only the ranks that own tokens < h*w see timestep 0
    assert [sp.first_frame_rows(880, 8, r, 27280) for r in range(8)] == [880, 0, 0, 0, 0, 0, 0, 0]
    assert [sp.first_frame_rows(390, 4, r, 300) for r in range(4)] == [75, 75, 75, 75]


class _FakeDit(torch.nn.Module):
    require_vae_embedding = False
    require_clip_embedding = False

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1), requires_grad=False)


def test_model_fn_rejects_out_of_scope_inputs():
    from fairygen_b200.model_fn import model_fn_wan_video

    lat = torch.zeros(1, 48, 2, 4, 4)
    for kw in ({"vace_context": object()}, {"audio_embeds": lat}, {"tea_cache": object()}, {"reference_latents": lat},
               {"sliding_window_size": 4}):
        with pytest.raises(NotImplementedError):
            model_fn_wan_video(dit=_FakeDit(), latents=lat, timestep=torch.ones(1), context=torch.zeros(1, 4, 8), **kw)
    with pytest.raises(ValueError):
        model_fn_wan_video(dit=_FakeDit(), latents=None, timestep=None, context=None)


def test_synthetic_state_dict_is_deterministic_and_lora_changes_targets_only():
    tiny = WanDiTConfig(dim=256, ffn_dim=512, text_dim=128, num_heads=2, num_layers=1)
    a = synthetic.random_state_dict(tiny, seed=0, device="cpu", dtype=torch.float32)
    b = synthetic.random_state_dict(tiny, seed=0, device="cpu", dtype=torch.float32, lora_rank=4)
    assert set(a) == set(synthetic.param_shapes(tiny))
    changed = {k for k in a if not torch.equal(a[k], b[k])}
    assert changed == {t + ".weight" for t in synthetic.lora_targets(tiny)}


def test_exchange_barrier_timeout_surfaces_on_the_host():
    """SequenceParallel.check reads the arena's status word (written by fgb_sp_barrier_status / fgb_sp_stats_barrier when a peer
    never reaches an epoch) and raises; WanDenoiser calls it once per video, after the loop."""
    import types

    from fairygen_b200.pipeline import WanDenoiser

    par = sp.SequenceParallel.__new__(sp.SequenceParallel)
    par.world, par.rank, par.exchange = 4, 2, "p2p"
    par.arena = None
    par.check()                                                     # nothing mapped yet: nothing to report
    par.arena = types.SimpleNamespace(status=torch.zeros(1, dtype=torch.int32), epoch=12)
    par.check()
    par.arena.status[0] = 9
    with pytest.raises(RuntimeError, match=r"rank 2 of 4 .* epoch 9 \(now at 12\)"):
        par.check()
    par.exchange = "nccl"                                           # the library variant has no flags of ours
    par.check()
    par.exchange = "p2p"

    den = WanDenoiser.__new__(WanDenoiser)
    den.engine = types.SimpleNamespace(device=torch.device("cpu"), sp=par)
    den._host_contexts = []
    den.scheduler = types.SimpleNamespace(timesteps=[])
    lat = torch.zeros(1, 4, 2, 2, 2)
    with pytest.raises(RuntimeError, match="never arrived"):
        den(lat, torch.zeros(1, 4, 8), None, steps=range(0))
    par.arena.status[0] = 0
    assert den(lat, torch.zeros(1, 4, 8), None, steps=range(0)).dtype == torch.bfloat16


def test_engine_for_refuses_a_dit_on_the_host():
    """The drop-in packs from device memory: a DiT whose weights still live on the host (the reference's offload modes) is an
    error with instructions at the first call, not a silent CPU path."""
    from fairygen_b200 import model_fn as mf

    dit = torch.nn.Linear(4, 4)
    with pytest.raises(RuntimeError, match="move the model to the GPU"):
        mf.engine_for(dit)
    mf._require_device_weights(torch.device("cuda", 0))      # a device pointer is all it asks for
