"""Host-side mirror of the reference 3D cache (Path R rows R5-R7 of SURVEY.md §8a), GPU resident.

reference: cosmos_predict1/diffusion/inference/cache_3d.py — Cache3D_Base :26-236, Cache3D_Buffer :239-343,
Cache3D_BufferSelector :346-420, Cache4D :423-433.  Same constructor keywords, ``render_cache`` / ``update_cache`` /
``input_frame_count`` signatures, output layouts and error behaviour.  Differences by design:
  * the cache lives in HBM (the reference parks it on the CPU and streams two frames per chunk over PCIe:
    cache_3d.py:97-101,183-223); one native call renders every target frame (`g3c_render_cache`), a second one applies the
    foreground-masking occlusion pass to all of them (`g3c_render_cache_occlusion`);
  * ``update_cache(depth_alignment=True)`` (the default) runs the non-rigid alignment as one native call
    (`g3c_align_depth_nonrigid`) instead of 100 autograd iterations.
"""
from __future__ import annotations

import torch

from . import camera_utils, warp

_AXES = ("B", "F", "N", "V", "C", "H", "W")


def _canonical(t: torch.Tensor, fmt) -> torch.Tensor:
    """Tensor whose dimensions are labelled by `fmt` -> 7-D view in the order B F N V C H W (missing axes have size 1)."""
    present = [a for a in _AXES if a in fmt]
    t = t.permute(*[list(fmt).index(a) for a in present])
    for i, a in enumerate(_AXES):
        if a not in fmt:
            t = t.unsqueeze(i)
    return t


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
        self.weight_dtype = weight_dtype
        self.is_depth = is_depth
        self.device = torch.device(device)
        self.filter_points_threshold = filter_points_threshold
        self.foreground_masking = foreground_masking
        if input_format is None:
            assert input_image.dim() == 4
            input_format = ["B", "C", "H", "W"]
        size = dict(zip(input_format, input_image.shape))
        B, Fr, N, V = (size.get(a, 1) for a in "BFNV")
        H, W = size.get("H"), size.get("W")
        image = _canonical(input_image, input_format).to(self.device, weight_dtype)
        # an explicit input mask travels as extra channels behind the 3 colour channels (reference :60-61,:97-99)
        self.input_mask = None if input_mask is None else _canonical(input_mask, input_format).to(self.device, weight_dtype)
        self.input_image = image[:, :, :, :, :3].contiguous() if input_mask is not None else image.contiguous()
        if input_points is not None:
            self.input_points = input_points.reshape(B, Fr, N, V, H, W, 3).to(self.device, weight_dtype)
            self.input_depth = None
        else:
            input_depth = torch.clamp(torch.nan_to_num(input_depth.to(self.device), nan=100), min=0, max=100)
            self.input_points = self._compute_input_points(
                input_depth.reshape(-1, 1, H, W), input_w2c.to(self.device).reshape(-1, 4, 4),
                input_intrinsics.to(self.device).reshape(-1, 3, 3)).to(weight_dtype).reshape(B, Fr, N, V, H, W, 3)
            self.input_depth = input_depth
        if self.filter_points_threshold < 1.0 and input_depth is not None:
            keep = warp.reliable_depth_mask_range_batch(input_depth.reshape(-1, 1, H, W),
                                                        ratio_thresh=self.filter_points_threshold)
            keep = keep.reshape(B, Fr, N, V, 1, H, W)
            self.input_mask = keep if self.input_mask is None else self.input_mask * keep.to(self.input_mask.device)
        self.boundary_mask = None
        if foreground_masking:  # depth-discontinuity pixels seed the occlusion mesh (reference :128-131)
            reliable = warp.reliable_depth_mask_range_batch(input_depth.reshape(-1, 1, H, W))
            self.boundary_mask = (~reliable).reshape(B, Fr, N, V, 1, H, W)

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
        if V != 1:
            raise NotImplementedError  # reference :229-230
        if Fr == 1:
            sl = slice(0, 1)
        else:
            sl = slice(start_frame_idx, start_frame_idx + F_target)
            if self.input_image[:, sl].shape[1] != F_target:
                raise RuntimeError(f"The expanded size of the tensor ({F_target}) must match the existing size "
                                   f"({self.input_image[:, sl].shape[1]}) at non-singleton dimension 1")  # torch .expand
        pts = self.input_points[:, sl, :, 0]
        img = self.input_image[:, sl, :, 0]
        msk = self.input_mask[:, sl, :, 0].to(torch.float32) if self.input_mask is not None else None
        bnd = None
        if self.foreground_masking:
            bsl = sl if self.boundary_mask.shape[1] != 1 else slice(0, 1)
            bnd = self.boundary_mask[:, bsl, :, 0, 0]
        return warp.render_cache(pts, img, msk, target_w2cs.to(self.device, torch.float32),
                                 target_intrinsics.to(self.device, torch.float32), render_depth=render_depth,
                                 boundary_masks=bnd)


class Cache3D_Buffer(Cache3D_Base):
    def __init__(self, frame_buffer_max=0, noise_aug_strength=0, generator=None, **kwargs):
        super().__init__(**kwargs)
        self.frame_buffer_max = frame_buffer_max
        self.noise_aug_strength = noise_aug_strength
        self.generator = generator

    def _insert_newest(self, name: str, new: torch.Tensor) -> None:
        """Slot 0 of the N axis is the newest entry.  While the ring is short the entry is prepended; once it holds
        frame_buffer_max entries slot 0 is overwritten and the older slots stay (reference :300-313)."""
        cur = getattr(self, name)
        entry = new[:, None, None, None].to(cur)
        if cur.shape[2] < self.frame_buffer_max:
            setattr(self, name, torch.cat([entry, cur], dim=2))
        else:
            cur[:, :, 0] = entry[:, :, 0]

    def update_cache(self, new_image, new_depth, new_w2c, new_mask=None, new_intrinsics=None, depth_alignment=True,
                     alignment_method="non_rigid"):
        """reference :246-316."""
        f32 = self.weight_dtype
        new_image = new_image.to(self.device, f32)
        new_depth = torch.clamp(torch.nan_to_num(new_depth.to(self.device, f32), nan=1e4), min=0, max=1e4)
        new_w2c = new_w2c.to(self.device, f32)
        if new_intrinsics is not None:
            new_intrinsics = new_intrinsics.to(self.device, f32)
        if depth_alignment:
            if alignment_method not in ("rigid", "non_rigid"):
                raise NotImplementedError
            target_depth, target_mask = self.render_cache(new_w2c.unsqueeze(1), new_intrinsics.unsqueeze(1),
                                                          render_depth=True)
            target_depth, target_mask = target_depth[:, :, 0], target_mask[:, :, 0]
            extra = {}
            if alignment_method == "non_rigid":
                extra = dict(k=new_intrinsics.squeeze(), c2w=torch.inverse(new_w2c.squeeze()),
                             alignment_method="non_rigid", num_iters=100, lambda_arap=0.1, smoothing_kernel_size=3)
            new_depth = camera_utils.align_depth(new_depth.squeeze(), target_depth.squeeze(),
                                                 target_mask.bool().squeeze(), **extra).reshape_as(new_depth)
        new_points = warp.unproject_points(new_depth, new_w2c, new_intrinsics, is_depth=self.is_depth)
        if self.filter_points_threshold < 1.0:
            B, Fr, N, V, C, H, W = self.input_image.shape
            keep = warp.reliable_depth_mask_range_batch(new_depth.reshape(-1, 1, H, W),
                                                        ratio_thresh=self.filter_points_threshold).reshape(B, 1, H, W)
            new_mask = keep if new_mask is None else new_mask.to(self.device) * keep
        if self.frame_buffer_max > 1:
            self._insert_newest("input_image", new_image)
            self._insert_newest("input_points", new_points)
            if self.input_mask is not None:
                self._insert_newest("input_mask", new_mask)
        else:  # a single-entry cache is replaced outright; the reference leaves the mask as it was (:314-316)
            self.input_image = new_image[:, None, None, None]
            self.input_points = new_points[:, None, None, None]

    def render_cache(self, target_w2cs, target_intrinsics, render_depth: bool = False, start_frame_idx: int = 0):
        assert start_frame_idx == 0, "start_frame_idx must be 0 for Cache3D_Buffer"
        output_device = target_w2cs.device
        pixels, masks = super().render_cache(target_w2cs, target_intrinsics, render_depth)
        pixels, masks = pixels.to(output_device), masks.to(output_device)
        if not render_depth:
            # reference :336-343: older buffers get more noise; the draw happens even at strength 0 (RNG stream parity)
            noise = torch.randn(pixels.shape, generator=self.generator, device=pixels.device, dtype=pixels.dtype)
            age = torch.arange(start=pixels.shape[2] - 1, end=-1, step=-1, device=pixels.device)
            pixels = pixels + noise * (age * self.noise_aug_strength).reshape(1, 1, -1, 1, 1, 1)
        return pixels, masks


class Cache3D_BufferSelector(Cache3D_Base):
    def __init__(self, frame_buffer_max=1, mask_for_max_buffer_model: bool = True, mask_full_threshold: float = 0.9,
                 **kwargs):
        """Many source frames on the N axis at construction; every render keeps the frame_buffer_max of them that cover
        the targets best (reference :346-420).  No update_cache."""
        super().__init__(**kwargs)
        self.frame_buffer_max = max(int(frame_buffer_max), 1)
        self.mask_for_max_buffer_model = bool(mask_for_max_buffer_model)
        self.mask_full_threshold = float(mask_full_threshold)

    def update_cache(self, *args, **kwargs):
        raise NotImplementedError("Cache3D_BufferSelector does not support update_cache")

    def render_cache(self, target_w2cs, target_intrinsics, render_depth: bool = False, start_frame_idx: int = 0):
        output_device = target_w2cs.device
        pixels, masks = super().render_cache(target_w2cs, target_intrinsics, render_depth, start_frame_idx)
        B, F, N = pixels.shape[:3]
        if N > self.frame_buffer_max:
            # overlap score of a buffer = covered pixels summed over target frames; keep the top-k per batch element
            score = masks.sum(dim=(1, 3, 4, 5))                                            # [B, N]
            top = score.topk(k=self.frame_buffer_max, dim=1, largest=True, sorted=True).indices
            pick = torch.arange(B, device=top.device)[:, None]
            pixels = pixels.transpose(1, 2)[pick, top].transpose(1, 2)                     # gather along N
            masks = masks.transpose(1, 2)[pick, top].transpose(1, 2)
        if self.mask_for_max_buffer_model and not render_depth:
            # per target frame keep ONE buffer — the first whose coverage reaches the threshold — or all of them if
            # none does (reference :397-418)
            cover = masks.mean(dim=[3, 4, 5])                                              # [B, F, k]
            full = cover >= self.mask_full_threshold
            first = torch.nn.functional.one_hot(full.float().argmax(dim=-1), cover.shape[-1]).to(cover.dtype)
            keep = torch.where(full.any(dim=-1, keepdim=True), first, torch.ones_like(cover))[..., None, None, None]
            pixels = (pixels + 1) * keep - 1
            masks = masks * keep
        return pixels.to(output_device), masks.to(output_device)


class Cache4D(Cache3D_Base):
    """One cache frame per target frame (dynamic scenes): `F` axis at construction, `start_frame_idx` at render
    (reference :423-433)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def update_cache(self, **kwargs):
        raise NotImplementedError

    def render_cache(self, target_w2cs, target_intrinsics, render_depth=False, start_frame_idx=0):
        return super().render_cache(target_w2cs, target_intrinsics, render_depth, start_frame_idx)
