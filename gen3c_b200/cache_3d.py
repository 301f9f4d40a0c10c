"""Host-side mirror of the reference 3D cache (Path R rows R5-R7 of SURVEY.md §8a), GPU resident.

reference: cosmos_predict1/diffusion/inference/cache_3d.py — Cache3D_Base :26-236,
Cache3D_Buffer :239-343.  Same constructor keywords, ``render_cache`` / ``update_cache`` /
``input_frame_count`` signatures, output layouts and error behaviour.  Differences by design:
  * the cache lives in HBM (the reference parks it on the CPU and streams two frames per chunk over
    PCIe: cache_3d.py:97-101,183-223); one native call renders every target frame;
  * ``foreground_masking`` (mesh occlusion pass) and ``update_cache(depth_alignment=True)`` are
    SURVEY.md §8(f) "next" rows and raise NotImplementedError until they are built.
"""
from __future__ import annotations

import torch

from . import warp


class Cache3D_Base:
    def __init__(
        self,
        input_image,
        input_depth,
        input_w2c,
        input_intrinsics,
        input_mask=None,
        input_format=None,
        input_points=None,
        weight_dtype=torch.float32,
        is_depth=True,
        device="cuda",
        filter_points_threshold=1.0,
        foreground_masking=False,
    ):
        """input_image: tensor whose dimensions are labelled by input_format, e.g. ['B','C','H','W'],
        ['B','N','C','H','W'], ['B','F','C','H','W'] (reference :41-45)."""
        if weight_dtype != torch.float32:
            raise NotImplementedError("the CUDA render path computes in float32 (reference default, cache_3d.py:36)")
        if foreground_masking:
            raise NotImplementedError("foreground_masking (mesh occlusion pass) is SURVEY.md §8(f) rank 1: not built yet")
        self.weight_dtype = weight_dtype
        self.is_depth = is_depth
        self.device = torch.device(device)
        self.filter_points_threshold = filter_points_threshold
        self.foreground_masking = foreground_masking
        if input_format is None:
            assert input_image.dim() == 4
            input_format = ["B", "C", "H", "W"]
        idx = {d: i for i, d in enumerate(input_format)}
        shape = input_image.shape
        if input_mask is not None:
            input_image = torch.cat([input_image, input_mask.to(input_image)], dim=idx.get("C"))
        B = shape[idx["B"]] if "B" in idx else 1
        Fr = shape[idx["F"]] if "F" in idx else 1
        N = shape[idx["N"]] if "N" in idx else 1
        V = shape[idx["V"]] if "V" in idx else 1
        H, W = shape[idx["H"]], shape[idx["W"]]
        if V != 1:
            raise NotImplementedError  # reference :229-230
        order = [idx.get(d) for d in ["B", "F", "N", "V", "C", "H", "W"]]
        input_image = input_image.permute(*[o for o in order if o is not None])
        for i, o in enumerate(order):
            if o is None:
                input_image = input_image.unsqueeze(i)
        input_image = input_image.to(self.device)
        if input_mask is not None:
            self.input_image, self.input_mask = input_image[:, :, :, :, :3], input_image[:, :, :, :, 3:]
        else:
            self.input_image, self.input_mask = input_image, None
        self.input_image = self.input_image.to(weight_dtype).contiguous()
        if input_points is not None:
            self.input_points = input_points.reshape(B, Fr, N, V, H, W, 3).to(self.device, weight_dtype)
            self.input_depth = None
        else:
            input_depth = torch.clamp(torch.nan_to_num(input_depth.to(self.device), nan=100), min=0, max=100)
            self.input_points = self._compute_input_points(
                input_depth.reshape(-1, 1, H, W), input_w2c.to(self.device).reshape(-1, 4, 4),
                input_intrinsics.to(self.device).reshape(-1, 3, 3)).reshape(B, Fr, N, V, H, W, 3)
            self.input_depth = input_depth
        if self.filter_points_threshold < 1.0 and input_depth is not None:
            dm = warp.reliable_depth_mask_range_batch(input_depth.reshape(-1, 1, H, W),
                                                      ratio_thresh=self.filter_points_threshold)
            dm = dm.reshape(B, Fr, N, V, 1, H, W)
            self.input_mask = dm if self.input_mask is None else self.input_mask * dm.to(self.input_mask)
        self.boundary_mask = None

    def _compute_input_points(self, input_depth, input_w2c, input_intrinsics):
        return warp.unproject_points(input_depth, input_w2c, input_intrinsics, is_depth=self.is_depth)

    def update_cache(self):
        raise NotImplementedError

    def input_frame_count(self) -> int:
        return self.input_image.shape[1]

    def render_cache(self, target_w2cs, target_intrinsics, render_depth=False, start_frame_idx=0):
        """reference :151-236 -> (pixels [B,F,N,3,H,W] or depth [B,F,N,H,W], masks [B,F,N,1,H,W])."""
        bs, F_target, _, _ = target_w2cs.shape
        B, Fr, N, V, C, H, W = self.input_image.shape
        assert bs == B
        if Fr == 1:
            sl = slice(0, 1)
        else:
            sl = slice(start_frame_idx, start_frame_idx + F_target)
            assert self.input_image[:, sl].shape[1] == F_target, "cache has fewer frames than targets"
        pts = self.input_points[:, sl, :, 0]
        img = self.input_image[:, sl, :, 0]
        msk = self.input_mask[:, sl, :, 0].to(torch.float32) if self.input_mask is not None else None
        return warp.render_cache(pts, img, msk, target_w2cs.to(self.device, torch.float32),
                                 target_intrinsics.to(self.device, torch.float32), render_depth=render_depth)


class Cache3D_Buffer(Cache3D_Base):
    def __init__(self, frame_buffer_max=0, noise_aug_strength=0, generator=None, **kwargs):
        super().__init__(**kwargs)
        self.frame_buffer_max = frame_buffer_max
        self.noise_aug_strength = noise_aug_strength
        self.generator = generator

    def update_cache(self, new_image, new_depth, new_w2c, new_mask=None, new_intrinsics=None, depth_alignment=True,
                     alignment_method="non_rigid"):
        """reference :246-316 (newest frame first in the N<=frame_buffer_max ring)."""
        if depth_alignment:
            raise NotImplementedError("depth alignment (100-step Adam, camera_utils.py:225-345) is SURVEY.md §8(f) rank 3")
        new_image = new_image.to(self.device, self.weight_dtype)
        new_depth = torch.clamp(torch.nan_to_num(new_depth.to(self.device, self.weight_dtype), nan=1e4), min=0, max=1e4)
        new_w2c = new_w2c.to(self.device, self.weight_dtype)
        new_intrinsics = new_intrinsics.to(self.device, self.weight_dtype)
        new_points = warp.unproject_points(new_depth, new_w2c, new_intrinsics, is_depth=self.is_depth)
        if self.filter_points_threshold < 1.0:
            B, Fr, N, V, C, H, W = self.input_image.shape
            dm = warp.reliable_depth_mask_range_batch(new_depth.reshape(-1, 1, H, W),
                                                      ratio_thresh=self.filter_points_threshold).reshape(B, 1, H, W)
            new_mask = dm if new_mask is None else new_mask.to(self.device) * dm
        if self.frame_buffer_max > 1:
            if self.input_image.shape[2] < self.frame_buffer_max:
                self.input_image = torch.cat([new_image[:, None, None, None], self.input_image], 2)
                self.input_points = torch.cat([new_points[:, None, None, None], self.input_points], 2)
                if self.input_mask is not None:
                    self.input_mask = torch.cat([new_mask[:, None, None, None].to(self.input_mask), self.input_mask], 2)
            else:
                self.input_image[:, :, 0] = new_image[:, None, None]
                self.input_points[:, :, 0] = new_points[:, None, None]
                if self.input_mask is not None:
                    self.input_mask[:, :, 0] = new_mask[:, None, None].to(self.input_mask)
        else:
            self.input_image = new_image[:, None, None, None]
            self.input_points = new_points[:, None, None, None]

    def render_cache(self, target_w2cs, target_intrinsics, render_depth: bool = False, start_frame_idx: int = 0):
        assert start_frame_idx == 0, "start_frame_idx must be 0 for Cache3D_Buffer"
        output_device = target_w2cs.device
        pixels, masks = super().render_cache(target_w2cs, target_intrinsics, render_depth)
        pixels, masks = pixels.to(output_device), masks.to(output_device)
        if not render_depth:
            # reference :336-343 (zero strength by default; the RNG draw is kept for stream compatibility)
            noise = torch.randn(pixels.shape, generator=self.generator, device=pixels.device, dtype=pixels.dtype)
            per_buffer = torch.arange(start=pixels.shape[2] - 1, end=-1, step=-1, device=pixels.device) * self.noise_aug_strength
            pixels = pixels + noise * per_buffer.reshape(1, 1, -1, 1, 1, 1)
        return pixels, masks
