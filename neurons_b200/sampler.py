"""Denoise-loop harness: the 25-step classifier-free-guidance DDIM loop that hosts the motion modules.

Mirrors the loop of NeuroclipsPipeline.__call__ (/root/reference/animatediff/pipelines/pipeline_neuroclips.py:433-483):
    for t in timesteps:  x2 = cat([latents]*2);  eps = denoiser(x2, t, ctx);  eps = eps_u + s*(eps_c - eps_u);  latents = ddim_step(...)
with the scheduler the reference configures (`DDIMScheduler(**noise_scheduler_kwargs)`, scripts/neuroclips_video_enhance.py:220;
configs/inference/inference-v3.yaml:16-21: linear betas 0.00085 -> 0.012, steps_offset 1, clip_sample false).  `DDIMScheduler`
lives in the un-vendored diffusers 0.11.1; its published algorithm (Song et al., DDIM, eta = 0) is restated here:
    a_t = prod(1 - beta)[t];   x0 = (x_t - sqrt(1 - a_t) eps) / sqrt(a_t);   x_{t'} = sqrt(a_{t'}) x0 + sqrt(1 - a_{t'}) eps
The denoiser is any callable -- in NEURONS the UNet3DConditionModel whose motion modules `neurons_b200.patch()` replaces; the
tests drive a stack of motion modules (`MotionStack`) so that the 25-step error accumulation of the CUDA path can be checked
against the CPU oracle without the reference UNet (which does not exist on the GPU box).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch


@dataclass
class DDIMSchedule:
    num_train_timesteps: int = 1000
    beta_start: float = 0.00085
    beta_end: float = 0.012
    steps_offset: int = 1
    set_alpha_to_one: bool = True

    def __post_init__(self):
        betas = torch.linspace(self.beta_start, self.beta_end, self.num_train_timesteps, dtype=torch.float32)   # "linear"
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).double()
        self.final_alpha_cumprod = 1.0 if self.set_alpha_to_one else float(self.alphas_cumprod[0])

    def timesteps(self, num_inference_steps: int) -> List[int]:
        """25 steps -> 961, 921, ..., 41, 1 (leading spacing + steps_offset)."""
        ratio = self.num_train_timesteps // num_inference_steps
        return [i * ratio + self.steps_offset for i in reversed(range(num_inference_steps))]

    def add_noise(self, x0: torch.Tensor, noise: torch.Tensor, t: int) -> torch.Tensor:
        a = float(self.alphas_cumprod[t])
        return (a ** 0.5) * x0 + ((1.0 - a) ** 0.5) * noise

    def alphas(self, t: int, num_inference_steps: int):
        """(alphas_cumprod[t], alphas_cumprod[previous timestep]) of one update."""
        prev_t = t - self.num_train_timesteps // num_inference_steps
        return float(self.alphas_cumprod[t]), float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else self.final_alpha_cumprod

    def step(self, eps: torch.Tensor, t: int, x: torch.Tensor, num_inference_steps: int) -> torch.Tensor:
        """Deterministic DDIM update (eta = 0, epsilon prediction, no sample clipping)."""
        a_t, a_prev = self.alphas(t, num_inference_steps)
        x0 = (x - ((1.0 - a_t) ** 0.5) * eps) / (a_t ** 0.5)
        return (a_prev ** 0.5) * x0 + ((1.0 - a_prev) ** 0.5) * eps


def denoise(denoiser: Callable[[torch.Tensor, int, Optional[torch.Tensor]], torch.Tensor], latents: torch.Tensor,
            context: Optional[torch.Tensor], schedule: DDIMSchedule, num_inference_steps: int = 25, guidance_scale: float = 8.5,
            noise: Optional[torch.Tensor] = None, low_strength: Optional[float] = None, fused_step: bool = False) -> torch.Tensor:
    """latents: [b, 4, f, h, w].  With `noise` + `low_strength` the clean latents are first noised to the start timestep exactly as
    pipeline_neuroclips.py:410-423 does; the loop then always runs over ALL timesteps (as the reference does, :433)."""
    ts = schedule.timesteps(num_inference_steps)
    if noise is not None:
        init = min(int(num_inference_steps * (low_strength if low_strength is not None else 1.0)), num_inference_steps)
        t_start = max(num_inference_steps - init, 0)
        steps = ts[:t_start]
        latents = schedule.add_noise(latents, noise, steps[0] if steps else ts[0])
    cfg = guidance_scale > 1.0
    if fused_step:
        latents = latents.contiguous().clone()            # updated in place by the fused kernel; never the caller's tensor
    with torch.no_grad():
        for t in ts:
            x2 = torch.cat([latents] * 2) if cfg else latents                     # :435
            eps = denoiser(x2, t, context).to(latents.dtype)                      # :470-475
            if fused_step:                                                        # :478-483 in one CUDA kernel (ops.cfg_ddim_step)
                from . import ops
                a_t, a_prev = schedule.alphas(t, num_inference_steps)
                eps_u, eps_c = eps.chunk(2) if cfg else (eps, None)
                latents = ops.cfg_ddim_step(latents, eps_u, eps_c, guidance_scale, a_t, a_prev)       # in place on our private copy
                continue
            if cfg:
                eps_u, eps_c = eps.chunk(2)                                       # :478-480
                eps = eps_u + guidance_scale * (eps_c - eps_u)
            latents = schedule.step(eps, t, latents, num_inference_steps)         # :483
    return latents


class MotionStack:
    """Synthetic denoiser made only of motion modules + parameter-free glue, shaped like the UNet's motion-module schedule
    (down levels, then up levels with skip connections).  `module_fn(i, x)` runs motion module i on x [b, c, f, h, w]; everything
    else (channel lifting by fixed random projections, 2x average pooling / nearest up-sampling, timestep / context injection) is
    plain torch shared verbatim by the CUDA path and the oracle path, so any difference comes from the motion modules."""

    def __init__(self, channels: Sequence[int], seed: int = 0, latent_channels: int = 4, context_dim: int = 16):
        self.channels = list(channels)                     # e.g. (64, 128): two modules per level down, two per level up
        g = torch.Generator().manual_seed(seed)
        self.latent_channels = latent_channels

        def rnd(o, i):
            return torch.randn(o, i, generator=g) / i ** 0.5
        self.lift = rnd(self.channels[0], latent_channels)
        self.down = [rnd(self.channels[k + 1], self.channels[k]) for k in range(len(self.channels) - 1)]
        self.up = [rnd(self.channels[k], self.channels[k + 1]) for k in range(len(self.channels) - 1)]
        self.out = rnd(latent_channels, self.channels[0]) * 0.1
        self.ctx = rnd(self.channels[0], context_dim)
        self.n_modules = 4 * len(self.channels) - 2 if len(self.channels) > 1 else 2

    def module_channels(self) -> List[int]:
        down = [c for c in self.channels for _ in range(2)]
        up = [c for c in reversed(self.channels[:-1]) for _ in range(2)]
        return down + up

    @staticmethod
    def _mix(w: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        return torch.einsum("oc,bcfhw->bofhw", w.to(x.device, x.dtype), x)

    def __call__(self, module_fn: Callable[[int, torch.Tensor], torch.Tensor], x: torch.Tensor, t: int,
                 context: Optional[torch.Tensor]) -> torch.Tensor:
        h = self._mix(self.lift, x)
        h = h + torch.tensor(t / 1000.0, dtype=h.dtype, device=h.device)
        if context is not None:                                                   # [b, context_dim] -> per-channel shift
            h = h + (context.to(h.device, h.dtype) @ self.ctx.to(h.device, h.dtype).T)[:, :, None, None, None]
        skips, i = [], 0
        for k, _ in enumerate(self.channels):
            for _ in range(2):
                h = module_fn(i, h); i += 1
            if k < len(self.channels) - 1:
                skips.append(h)
                b, c, f, hh, ww = h.shape
                h = h.reshape(b, c, f, hh // 2, 2, ww // 2, 2).mean(dim=(4, 6))   # 2x average pool
                h = self._mix(self.down[k], h)
        for k in reversed(range(len(self.channels) - 1)):
            h = self._mix(self.up[k], h)
            h = h.repeat_interleave(2, dim=3).repeat_interleave(2, dim=4) + skips.pop()
            for _ in range(2):
                h = module_fn(i, h); i += 1
        return self._mix(self.out, h)
