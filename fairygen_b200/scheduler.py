"""Flow-matching scheduler for the "Wan" template with the update fused into one sm_100a kernel.

API-compatible with the reference ``FlowMatchScheduler`` for the methods the hot path uses
(animation/diffsynth/diffusion/flow_match.py:29-39 set_timesteps_wan, :132-142 set_timesteps,
:144-154 step, :164-179 add_noise / training_target / training_weight).  The schedule itself is a
few dozen host floats, built with the same fp32 torch ops as the reference so sigmas/timesteps are
bit-identical; the per-step tensor update runs in ``fgb_cfg_fm_step``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class FlowMatchScheduler:
    def __init__(self, template: str = "Wan"):
        if template != "Wan":
            raise NotImplementedError(f"template {template!r}: only the 'Wan' schedule is on the hot path")
        self.num_train_timesteps = 1000
        self.sigmas = None
        self.timesteps = None
        self.training = False

    @staticmethod
    def set_timesteps_wan(num_inference_steps=100, denoising_strength=1.0, shift=None):
        shift = 5 if shift is None else shift
        sigma_start = 0.0 + (1.0 - 0.0) * denoising_strength
        sigmas = torch.linspace(sigma_start, 0.0, num_inference_steps + 1)[:-1]
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        return sigmas, sigmas * 1000

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False, **kwargs):
        self.sigmas, self.timesteps = self.set_timesteps_wan(num_inference_steps, denoising_strength, **kwargs)
        self.training = bool(training)
        if training:
            steps = 1000
            y = torch.exp(-2 * ((self.timesteps - steps / 2) / steps) ** 2)
            y = y - y.min()
            wgt = y * (steps / y.sum())
            if len(self.timesteps) != 1000:
                wgt = wgt * (len(self.timesteps) / steps)
                wgt = wgt + wgt[1]
            self.linear_timesteps_weights = wgt

    # ---- helpers ----------------------------------------------------------------------------------
    def _index(self, timestep) -> int:
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.detach().cpu()
        return int(torch.argmin((self.timesteps - timestep).abs()))

    def sigma_delta(self, index: int, to_final: bool = False) -> float:
        """fp32 (sigma_next - sigma) exactly as the reference forms it (FM:148-153)."""
        sigma = self.sigmas[index]
        if to_final or index + 1 >= len(self.timesteps):
            return float(0 - sigma)
        return float(self.sigmas[index + 1] - sigma)

    # ---- reference-compatible step (returns a new tensor, FM:144-154) -----------------------------
    def step(self, model_output, timestep, sample, to_final=False, **kwargs):
        out = sample.clone(memory_format=torch.contiguous_format)
        self.step_fused(out, model_output, None, 1.0, self._index(timestep), None, to_final=to_final)
        return out

    # ---- fused: CFG combine + Euler update + first-frame restore, in place (PIPE:302-309) ----------
    def step_fused(self, latents, noise_pos, noise_neg: Optional[torch.Tensor], cfg_scale: float, index: int,
                   first_frame_latents: Optional[torch.Tensor] = None, to_final: bool = False):
        if latents.dtype != torch.bfloat16 or not latents.is_cuda:
            raise NotImplementedError("fairygen_b200 scheduler: the fused step runs on bf16 CUDA latents only (no fallback)")
        c = lambda t: None if t is None else t.to(dtype=torch.bfloat16).contiguous()  # noqa: E731
        ops.cfg_fm_step(latents, c(noise_pos), c(noise_neg), c(first_frame_latents), float(cfg_scale),
                        self.sigma_delta(index, to_final))
        return latents

    # ---- training-side helpers (elementwise, host-agnostic; FM:164-179) -----------------------------
    def add_noise(self, original_samples, noise, timestep):
        sigma = self.sigmas[self._index(timestep)]
        return (1 - sigma) * original_samples + sigma * noise

    def training_target(self, sample, noise, timestep):
        return noise - sample

    def training_weight(self, timestep):
        return self.linear_timesteps_weights[self._index(timestep)]
