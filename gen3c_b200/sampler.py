"""Host-side mirror of the reference sampler loop (Path D, rows D1/D2 of SURVEY.md §8a).

reference: cosmos_predict1/diffusion/model/model_v2w.py:84-155 (generate_samples_from_batch),
:201-259 (_augment_noise_with_latent, _reverse_precondition_*), utils/misc.py:133-154
(arch_invariant_rand) and diffusers 0.32.2 EDMEulerScheduler (Karras sigmas, Euler step; restated —
the package is not in the reference tree).  The loop body runs as ONE native call
(``g3c_denoise_step``): sampler glue + two DiT forwards; the constant augmentation noise is drawn
once instead of every step (the reference re-draws the same numbers: model_v2w.py:232-237).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from .dit import VideoExtendGeneralDIT


class EDMEulerScheduler:
    """The subset of diffusers' EDMEulerScheduler the reference touches (model_t2w.py:65, model_v2w.py:121-149)."""

    def __init__(self, sigma_max: float = 80.0, sigma_min: float = 0.0002, sigma_data: float = 0.5, rho: float = 7.0):
        self.sigma_max, self.sigma_min, self.sigma_data, self.rho = sigma_max, sigma_min, sigma_data, rho
        self.sigmas = None
        self.timesteps = None

    @property
    def init_noise_sigma(self) -> float:
        return math.sqrt(self.sigma_max ** 2 + 1)

    def set_timesteps(self, num_steps: int):
        ramp = np.linspace(0, 1, num_steps)
        lo, hi = self.sigma_min ** (1 / self.rho), self.sigma_max ** (1 / self.rho)
        sig = (hi + ramp * (lo - hi)) ** self.rho
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = 0.25 * np.log(self.sigmas[:-1])
        return self


def arch_invariant_rand(shape, seed: int) -> torch.Tensor:
    """utils/misc.py:133-154 — numpy RandomState(seed).standard_normal(shape), float32 (host)."""
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def denoise_step(net: VideoExtendGeneralDIT, xt: torch.Tensor, gt_latent: torch.Tensor, aug_noise: torch.Tensor,
                 indicator: torch.Tensor, cond_mask: torch.Tensor, pose_cond: Optional[torch.Tensor],
                 padding_mask: Optional[torch.Tensor], ctx_cond: torch.Tensor, ctx_uncond: torch.Tensor, sigma: float,
                 sigma_next: float, guidance: float, sigma_data: float = 0.5, sigma_aug: float = 0.001,
                 fps: float = 24.0, out: Optional[torch.Tensor] = None,
                 cond_mask_uncond: Optional[torch.Tensor] = None,
                 net_output: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One loop body of generate_samples_from_batch for B = 1 (this rank's T slice under CP).
    xt, gt_latent bf16 [16,T,H,W]; aug_noise f32 [16,T,H,W]; indicator f32 [T]; cond_mask bf16 [1,T,H,W];
    pose_cond bf16 [64,T,H,W]; padding_mask bf16 [H,W] (latent resolution) or None; ctx_* bf16 [M, ctx_dim].
    cond_mask_uncond: uncondition.condition_video_input_mask when it differs from the conditional one
    (add_input_frames_guidance, model_v2w.py:76-80); net_output: optional bf16 [16,T,H,W] receiving the CFG-combined
    network output of model_v2w.py:143."""
    _, T, H, W = xt.shape
    net._sync_weights()
    net._set_shape(T, H, W, ctx_cond.shape[0], fps)
    if out is None:
        out = torch.empty_like(xt)
    a = _lib.StepArgs(_lib.ptr(xt), _lib.ptr(gt_latent), _lib.ptr(aug_noise), _lib.ptr(indicator), _lib.ptr(cond_mask),
                      _lib.ptr(pose_cond), _lib.ptr(padding_mask), _lib.ptr(ctx_cond), _lib.ptr(ctx_uncond),
                      sigma, sigma_next, sigma_data, sigma_aug, guidance, _lib.ptr(out), _lib.ptr(cond_mask_uncond),
                      _lib.ptr(net_output))
    with torch.cuda.device(xt.device):
        _lib.check(_lib.load().g3c_denoise_step(net._engine(), C.byref(a), _lib.stream_ptr()), "g3c_denoise_step")
    return out


@torch.no_grad()
def generate_samples(net: VideoExtendGeneralDIT, state_shape, gt_latent: torch.Tensor, indicator: torch.Tensor,
                     cond_mask: torch.Tensor, pose_cond: torch.Tensor, padding_mask: Optional[torch.Tensor],
                     ctx_cond: torch.Tensor, ctx_uncond: torch.Tensor, guidance: float = 1.0, seed: int = 1,
                     num_steps: int = 35, sigma_aug: float = 0.001, fps: float = 24.0) -> torch.Tensor:
    """The sampler loop of generate_samples_from_batch for B = 1 without context parallelism.
    All condition tensors are full-T bf16 CUDA tensors (layouts as in `denoise_step`)."""
    sch = EDMEulerScheduler().set_timesteps(num_steps)
    dev = gt_latent.device
    g = torch.Generator(device=dev).manual_seed(seed)
    xt = (torch.randn(tuple(state_shape), device=dev, dtype=torch.bfloat16, generator=g) * sch.init_noise_sigma)
    noise = arch_invariant_rand(tuple(state_shape), seed).to(dev)
    for i in range(num_steps):
        xt = denoise_step(net, xt, gt_latent, noise, indicator, cond_mask, pose_cond, padding_mask, ctx_cond,
                          ctx_uncond, float(sch.sigmas[i]), float(sch.sigmas[i + 1]), guidance, sch.sigma_data,
                          sigma_aug, fps)
    return xt
