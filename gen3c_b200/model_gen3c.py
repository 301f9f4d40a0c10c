"""Host-side mirror of the reference diffusion model wrapper (Path D rows D1, D11, D12 of SURVEY.md §8a).

reference: cosmos_predict1/diffusion/model/model_gen3c.py :32-139 (encode_warped_frames, _get_conditions,
add_condition_pose), model/model_v2w.py :32-82 (add_condition_video_indicator_and_video_input_mask), :84-155
(generate_samples_from_batch), module/parallel.py (split / gather along latent T).

The tokenizer VAE and the T5 encoder are outside the hot path: they are injected as
callables, so the same code runs with the real TorchScript VAE or with a synthetic stand-in in tests.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional

import torch

from . import sampler
from .dit import VideoExtendGeneralDIT
from .parallel import cat_outputs_cp, chunk_bounds


@dataclass
class VideoExtendCondition:
    """The fields of the reference's VideoExtendCondition (conditioner.py:107-134) that reach the network."""

    crossattn_emb: torch.Tensor                              # [B, 512, 1024]
    padding_mask: Optional[torch.Tensor] = None              # [B, 1, H_pix, W_pix]
    fps: Optional[torch.Tensor] = None
    video_cond_bool: Optional[bool] = None
    gt_latent: Optional[torch.Tensor] = None                 # [B, 16, T, H, W]
    condition_video_indicator: Optional[torch.Tensor] = None  # [1, 1, T, 1, 1]
    condition_video_input_mask: Optional[torch.Tensor] = None  # [B, 1, T, H, W]
    condition_video_pose: Optional[torch.Tensor] = None       # [B, 64, T, H, W]
    extra: dict = field(default_factory=dict)

    def to_dict(self) -> dict:
        d = {k: getattr(self, k) for k in ("crossattn_emb", "padding_mask", "fps", "video_cond_bool", "gt_latent",
                                           "condition_video_indicator", "condition_video_input_mask",
                                           "condition_video_pose")}
        d.update(self.extra)
        return d


def encode_warped_frames(condition_state: torch.Tensor, condition_state_mask: torch.Tensor,
                         encode: Callable[[torch.Tensor], torch.Tensor], frame_buffer_max: int = 2,
                         dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """reference model_gen3c.py:32-57.  condition_state [B, F, N, 3, H, W] (warped frames), condition_state_mask
    [B, F, N, 1, H, W]; `encode` maps [B, 3, F, H, W] -> latent [B, 16, T, h, w].  Returns [B, 32*frame_buffer_max, T, h, w]."""
    assert condition_state.dim() == 6
    condition_state_mask = (condition_state_mask * 2 - 1).repeat(1, 1, 1, 3, 1, 1)
    latent_condition = []
    for i in range(condition_state.shape[2]):
        cur_video = encode(condition_state[:, :, i].permute(0, 2, 1, 3, 4).to(dtype)).contiguous()
        cur_mask = encode(condition_state_mask[:, :, i].permute(0, 2, 1, 3, 4).to(dtype)).contiguous()
        latent_condition += [cur_video, cur_mask]
    for _ in range(frame_buffer_max - condition_state.shape[2]):
        latent_condition += [torch.zeros_like(cur_video), torch.zeros_like(cur_mask)]
    return torch.cat(latent_condition, dim=1)


def add_condition_video_indicator_and_video_input_mask(latent_state: torch.Tensor, condition: VideoExtendCondition,
                                                       num_condition_t: Optional[int] = None) -> VideoExtendCondition:
    """reference model_v2w.py:32-82 (inference branch: condition_location = first_n)."""
    T = latent_state.shape[2]
    ind = torch.zeros(1, 1, T, 1, 1, device=latent_state.device).type(latent_state.dtype)
    assert num_condition_t is not None, "num_condition_t should be provided"
    assert num_condition_t <= T, f"num_condition_t should be less than T, get {num_condition_t}, {T}"
    ind[:, :, :num_condition_t] += 1.0
    condition.gt_latent = latent_state
    condition.condition_video_indicator = ind
    B, C, T, H, W = latent_state.shape
    ones = torch.ones((B, 1, T, H, W), dtype=latent_state.dtype, device=latent_state.device)
    zeros = torch.zeros_like(ones)
    assert condition.video_cond_bool is not None, "video_cond_bool should be set"
    condition.condition_video_input_mask = ind * ones + (1 - ind) * zeros if condition.video_cond_bool else zeros
    return condition


def add_condition_pose(latent_condition: torch.Tensor, condition: VideoExtendCondition,
                       drop_out_latent: bool = False) -> VideoExtendCondition:
    """reference model_gen3c.py:115-139 (the broadcast over the CP group is the caller's job here: every rank of
    the reference already computes identical conditions, gen3c_single_image.py:283-477)."""
    condition.condition_video_pose = (torch.zeros_like(latent_condition) if drop_out_latent
                                      else latent_condition).contiguous()
    return condition


def get_conditions(crossattn_emb: torch.Tensor, negative_crossattn_emb: torch.Tensor, padding_mask: torch.Tensor,
                   condition_state: torch.Tensor, condition_state_mask: torch.Tensor, condition_latent: torch.Tensor,
                   num_condition_t: int, encode: Callable[[torch.Tensor], torch.Tensor], frame_buffer_max: int = 2,
                   add_input_frames_guidance: bool = False, fps: Optional[torch.Tensor] = None,
                   dtype: torch.dtype = torch.bfloat16):
    """reference model_gen3c.py:59-113 (_get_conditions, negative-prompt branch)."""
    cond = VideoExtendCondition(crossattn_emb=crossattn_emb, padding_mask=padding_mask, fps=fps)
    uncond = VideoExtendCondition(crossattn_emb=negative_crossattn_emb, padding_mask=padding_mask, fps=fps)
    latent_condition = encode_warped_frames(condition_state, condition_state_mask, encode, frame_buffer_max, dtype)
    cond.video_cond_bool = True
    cond = add_condition_video_indicator_and_video_input_mask(condition_latent, cond, num_condition_t)
    cond = add_condition_pose(latent_condition, cond)
    uncond.video_cond_bool = False if add_input_frames_guidance else True
    uncond = add_condition_video_indicator_and_video_input_mask(condition_latent, uncond, num_condition_t)
    uncond = add_condition_pose(latent_condition, uncond, drop_out_latent=True)
    assert cond.gt_latent.allclose(uncond.gt_latent)
    return cond, uncond


@torch.no_grad()
def generate_samples_from_batch(net: VideoExtendGeneralDIT, condition: VideoExtendCondition,
                                uncondition: VideoExtendCondition, guidance: float = 1.5, seed: int = 1,
                                state_shape=(16, 16, 88, 160), num_steps: int = 35,
                                condition_augment_sigma: float = 0.001, sigma_data: float = 0.5,
                                xt0: Optional[torch.Tensor] = None) -> torch.Tensor:
    """reference model_v2w.py:84-155 for n_sample = 1: initial noise, 35 x loop body, CP split of the latent along T
    and the final all-gather.  The loop body is one native call (`g3c_denoise_step`)."""
    dev = condition.gt_latent.device
    bf = torch.bfloat16
    sch = sampler.EDMEulerScheduler(sigma_data=sigma_data).set_timesteps(num_steps)
    if xt0 is None:  # reference: torch.randn(...) * scheduler.init_noise_sigma (model_v2w.py:124)
        xt = torch.randn((1,) + tuple(state_shape), device=dev, dtype=bf) * sch.init_noise_sigma
    else:
        xt = xt0.to(dev, bf)
    noise = sampler.arch_invariant_rand((1,) + tuple(state_shape), seed).to(dev)
    T = state_shape[1]
    if net.is_context_parallel_enabled:
        start, length = chunk_bounds(T, net._cp_rank, net._cp_size)
    else:
        start, length = 0, T
    sl = slice(start, start + length)

    def loc(t):  # this rank's slice along latent T of a [B, C, T, H, W] tensor, batch element 0
        return t[0, :, sl].to(bf).contiguous()

    H, W = state_shape[2], state_shape[3]
    pm = condition.padding_mask
    pad = None
    if pm is not None:
        pad = torch.nn.functional.interpolate(pm.float(), size=(H, W), mode="nearest")[0, 0].to(bf).contiguous()
    fps = float(condition.fps.flatten()[0]) if condition.fps is not None else 24.0
    x = loc(xt)
    gt, mask, pose = loc(condition.gt_latent), loc(condition.condition_video_input_mask), loc(condition.condition_video_pose)
    # the unconditional branch sees uncondition's own input mask (all zeros under add_input_frames_guidance,
    # model_v2w.py:76-80) and a zero pose (model_gen3c.py:101-104, drop_out_latent=True)
    mask_u = loc(uncondition.condition_video_input_mask)
    if uncondition.condition_video_pose is not None and bool(uncondition.condition_video_pose.any()):
        raise NotImplementedError("the unconditional branch of GEN3C drops the pose latents (model_gen3c.py:103); "
                                  "a non-zero uncondition.condition_video_pose is not supported")
    ind = condition.condition_video_indicator[0, 0, sl, 0, 0].float().contiguous()
    nz = noise[0, :, sl].contiguous()
    ctx_c = condition.crossattn_emb[0].to(bf).contiguous()
    ctx_u = uncondition.crossattn_emb[0].to(bf).contiguous()
    for i in range(num_steps):
        x = sampler.denoise_step(net, x, gt, nz, ind, mask, pose, pad, ctx_c, ctx_u, float(sch.sigmas[i]),
                                 float(sch.sigmas[i + 1]), guidance, sigma_data, condition_augment_sigma, fps,
                                 cond_mask_uncond=mask_u)
    samples = x[None]
    if net.is_context_parallel_enabled:
        samples = cat_outputs_cp(samples, seq_dim=2, cp_group=net.cp_group)
    return samples


class DiffusionGen3CModel:
    """The reference's model object (model_gen3c.py:26-139 over model_v2w.py / model_t2w.py) reduced to what inference
    touches: `net`, `tokenizer`, `encode` / `decode`, `state_shape`, `frame_buffer_max`, `chunk_size`,
    `_get_conditions` and `generate_samples_from_batch` with the reference's signatures.  GEN3C_Cosmos_7B is the one
    configuration (config/inference/cosmos-1-diffusion-gen3c.py:22-46): frame_buffer_max 2, sigma_data 0.5, latent
    [16, 16, 88, 160], 121-frame tokenizer chunks."""

    def __init__(self, net: Optional[VideoExtendGeneralDIT] = None, tokenizer=None, frame_buffer_max: int = 2,
                 sigma_data: float = 0.5, state_shape=(16, 16, 88, 160), device="cuda"):
        self.device = torch.device(device)
        self.net = net
        self.tokenizer = tokenizer
        self.frame_buffer_max = frame_buffer_max
        self.chunk_size = 121
        self.sigma_data = sigma_data
        self.state_shape = list(state_shape)
        self.tensor_kwargs = {"device": self.device, "dtype": torch.bfloat16}
        self.cp_group = None   # set by the caller together with net.enable_context_parallel

    # model_t2w.py: encode / decode scale by sigma_data
    @torch.no_grad()
    def encode(self, state: torch.Tensor) -> torch.Tensor:
        return self.tokenizer.encode(state) * self.sigma_data

    @torch.no_grad()
    def decode(self, latent: torch.Tensor) -> torch.Tensor:
        return self.tokenizer.decode(latent / self.sigma_data)

    def encode_warped_frames(self, condition_state, condition_state_mask, dtype):
        return encode_warped_frames(condition_state, condition_state_mask, self.encode, self.frame_buffer_max, dtype)

    def _get_conditions(self, data_batch: dict, is_negative_prompt: bool = False,
                        condition_latent: Optional[torch.Tensor] = None, num_condition_t: Optional[int] = None,
                        add_input_frames_guidance: bool = False):
        """reference :59-113.  Without a negative prompt the unconditional text context is zeros (the conditioner's
        dropout of the text embedding, conditioner.py get_condition_uncondition)."""
        ctx = data_batch["t5_text_embeddings"]
        if is_negative_prompt:  # conditioner.py:267-292: falls back to the prompt itself when no negative embedding is given
            neg = data_batch.get("neg_t5_text_embeddings", ctx)
        else:                   # conditioner.py:234-265: the text embedder is dropped out (zeros)
            neg = torch.zeros_like(ctx)
        cond, uncond = get_conditions(ctx, neg, data_batch.get("padding_mask"), data_batch["condition_state"],
                                      data_batch["condition_state_mask"], condition_latent, num_condition_t, self.encode,
                                      self.frame_buffer_max, add_input_frames_guidance, data_batch.get("fps"),
                                      self.tensor_kwargs["dtype"])
        if self.net.is_context_parallel_enabled:
            from .parallel import broadcast_condition

            cond = broadcast_condition(cond, cp_group=self.net.cp_group)
            uncond = broadcast_condition(uncond, cp_group=self.net.cp_group)
        return cond, uncond

    @torch.no_grad()
    def generate_samples_from_batch(self, data_batch: dict, guidance: float = 1.5, seed: int = 1, state_shape=None,
                                    n_sample: Optional[int] = 1, is_negative_prompt: bool = False, num_steps: int = 35,
                                    condition_latent: Optional[torch.Tensor] = None,
                                    num_condition_t: Optional[int] = None, condition_augment_sigma: float = None,
                                    add_input_frames_guidance: bool = False) -> torch.Tensor:
        """reference model_v2w.py:84-155."""
        assert condition_latent is not None, "condition_latent should be provided"
        if n_sample not in (None, 1):
            raise NotImplementedError("the native loop body handles one sample per call (GEN3C inference uses n_sample=1)")
        cond, uncond = self._get_conditions(data_batch, is_negative_prompt, condition_latent, num_condition_t,
                                            add_input_frames_guidance)
        return generate_samples_from_batch(self.net, cond, uncond, guidance=guidance, seed=seed,
                                           state_shape=tuple(state_shape or self.state_shape), num_steps=num_steps,
                                           condition_augment_sigma=condition_augment_sigma, sigma_data=self.sigma_data)
