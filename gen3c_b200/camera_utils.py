"""Host-side mirror of the reference's camera helpers (the callers either side of Path R).

reference: cosmos_predict1/diffusion/inference/camera_utils.py — look_at_matrix :30-46, create_horizontal_trajectory
:48-88, create_spiral_trajectory :91-139, generate_camera_trajectory :142-222, _align_inv_depth_to_depth :225-270,
align_depth :273-347.  Same function names, arguments and error behaviour.  The trajectory builders are 4x4 host
arithmetic (vectorised here over the steps instead of a Python loop of tiny tensors); `align_depth` runs on the GPU: the
rigid stage is a quantile + 2-parameter least-squares fit, the non-rigid stage — 100 Adam steps that the reference drives
through autograd, ~4 000 ATen launches — is ONE native call (`g3c_align_depth_nonrigid`, closed-form gradient).
"""
from __future__ import annotations

import math

import torch

from . import _lib


def apply_transformation(Bx4x4: torch.Tensor, another_matrix: torch.Tensor) -> torch.Tensor:
    if another_matrix.dim() == 2:
        another_matrix = another_matrix.unsqueeze(0).expand(Bx4x4.shape[0], -1, -1)
    return torch.bmm(Bx4x4, another_matrix)


def _look_at_batch(camera_pos: torch.Tensor, target: torch.Tensor, invert_pos: bool = True) -> torch.Tensor:
    """[n,3] camera positions and look-at points -> [n,4,4]; rows right / up / forward, Y-up world."""
    forward = (target - camera_pos).float()
    forward = forward / forward.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 1.0, 0.0], device=camera_pos.device).expand_as(forward)
    right = torch.cross(up, forward, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    up = torch.cross(forward, right, dim=-1)
    m = torch.eye(4, device=camera_pos.device).repeat(camera_pos.shape[0], 1, 1)
    m[:, 0, :3], m[:, 1, :3], m[:, 2, :3] = right, up, forward
    m[:, :3, 3] = -camera_pos if invert_pos else camera_pos
    return m


def look_at_matrix(camera_pos: torch.Tensor, target: torch.Tensor, invert_pos: bool = True) -> torch.Tensor:
    return _look_at_batch(camera_pos[None].float(), target[None].float(), invert_pos)[0]


def _aim(look_at: torch.Tensor, pos: torch.Tensor, camera_rotation: str, allowed: str) -> torch.Tensor:
    if camera_rotation == "trajectory_aligned":
        return look_at + pos * 2
    if camera_rotation == "center_facing":
        return look_at.expand_as(pos)
    if camera_rotation == "no_rotation":
        return look_at + pos
    raise ValueError(f"Camera rotation should be {allowed}")


def create_horizontal_trajectory(world_to_camera_matrix, center_depth, positive=True, n_steps=13, distance=0.1,
                                 device="cuda", axis="x", camera_rotation="center_facing"):
    if axis not in ("x", "y", "z"):
        raise ValueError("Axis should be x, y or z")
    look_at = torch.tensor([0.0, 0.0, center_depth], device=device)
    # offsets i * distance * center_depth / n_steps, evaluated in Python floats like the reference's loop (:55-72)
    off = torch.tensor([i * distance * center_depth / n_steps * (1 if positive else -1) for i in range(n_steps)],
                       device=device)
    pos = torch.zeros(n_steps, 3, device=device)
    pos[:, "xyz".index(axis)] = off
    traj = _look_at_batch(pos, _aim(look_at, pos, camera_rotation, "center_facing or trajectory_aligned"))
    return apply_transformation(traj, world_to_camera_matrix)


def create_spiral_trajectory(world_to_camera_matrix, center_depth, radius_x=0.03, radius_y=0.02, radius_z=0.0,
                             positive=True, camera_rotation="center_facing", n_steps=13, device="cuda",
                             start_from_zero=True, num_circles=1):
    look_at = torch.tensor([0.0, 0.0, center_depth], device=device)
    theta_max = 2 * math.pi * num_circles
    rows = []
    for i in range(n_steps):
        theta = theta_max * i / (n_steps - 1)
        if start_from_zero:
            x = radius_x * (math.cos(theta) - 1) * (1 if positive else -1) * center_depth
        else:
            x = radius_x * math.cos(theta) * center_depth
        rows.append([x, radius_y * math.sin(theta) * center_depth, radius_z * math.sin(theta) * center_depth])
    pos = torch.tensor(rows, device=device)
    traj = _look_at_batch(pos, _aim(look_at, pos, camera_rotation, "center_facing, trajectory_aligned or no_rotation"))
    return apply_transformation(traj, world_to_camera_matrix)


_DIRECTIONS = {"left": (False, "x"), "right": (True, "x"), "up": (False, "y"), "down": (True, "y"),
               "zoom_in": (True, "z"), "zoom_out": (False, "z")}


def generate_camera_trajectory(trajectory_type: str, initial_w2c: torch.Tensor, initial_intrinsics: torch.Tensor,
                               num_frames: int, movement_distance: float, camera_rotation: str,
                               center_depth: float = 1.0, device: str = "cuda"):
    """-> (w2cs [1, num_frames, 4, 4], intrinsics [1, num_frames, 3, 3])."""
    if trajectory_type in ("clockwise", "counterclockwise"):
        seq = create_spiral_trajectory(world_to_camera_matrix=initial_w2c, center_depth=center_depth, n_steps=num_frames,
                                       positive=trajectory_type == "clockwise", device=device,
                                       camera_rotation=camera_rotation, radius_x=movement_distance,
                                       radius_y=movement_distance)
    elif trajectory_type in _DIRECTIONS:
        positive, axis = _DIRECTIONS[trajectory_type]
        seq = create_horizontal_trajectory(world_to_camera_matrix=initial_w2c, center_depth=center_depth,
                                           n_steps=num_frames, positive=positive, axis=axis,
                                           distance=movement_distance, device=device, camera_rotation=camera_rotation)
    else:
        raise ValueError(f"Unsupported trajectory type: {trajectory_type}")
    w2cs = seq.unsqueeze(0)
    if initial_intrinsics.dim() == 2:
        Ks = initial_intrinsics.unsqueeze(0).unsqueeze(0).repeat(1, num_frames, 1, 1)
    else:
        Ks = initial_intrinsics.unsqueeze(0)
    return w2cs, Ks


# ------------------------------------------------------------------------------------------------------
# depth alignment (update_cache)
# ------------------------------------------------------------------------------------------------------
def _align_inv_depth_to_depth(source_inv_depth: torch.Tensor, target_depth: torch.Tensor,
                              target_mask: torch.Tensor | None = None) -> torch.Tensor:
    """reference :225-270 — (h, w) tensors on one device; returns the aligned depth."""
    target_inv_depth = 1.0 / target_depth
    source_mask = source_inv_depth > 0
    target_depth_mask = target_depth > 0
    target_mask = target_depth_mask if target_mask is None else torch.logical_and(target_mask > 0, target_depth_mask)
    q = torch.tensor([0.1, 0.9], device=source_inv_depth.device)
    s_lo, s_hi = torch.quantile(source_inv_depth[source_mask], q)
    t_lo, t_hi = torch.quantile(target_inv_depth[target_mask], q)
    keep = ((source_inv_depth > s_lo) & (source_inv_depth < s_hi) &
            (target_inv_depth > t_lo) & (target_inv_depth < t_hi))
    # two-parameter least squares  [s 1] [scale bias]^T = t  through the normal equations in float64
    s, t = source_inv_depth[keep].double(), target_inv_depth[keep].double()
    n = s.numel()
    sx, sy, sxx, sxy = s.sum(), t.sum(), (s * s).sum(), (s * t).sum()
    det = n * sxx - sx * sx
    scale = ((n * sxy - sx * sy) / det).to(source_inv_depth.dtype)
    bias = ((sxx * sy - sx * sxy) / det).to(source_inv_depth.dtype)
    return 1.0 / (source_inv_depth * scale + bias)


def align_depth(source_depth: torch.Tensor, target_depth: torch.Tensor, target_mask: torch.Tensor,
                k: torch.Tensor = None, c2w: torch.Tensor = None, alignment_method: str = "rigid", num_iters: int = 100,
                lambda_arap: float = 0.1, smoothing_kernel_size: int = 3) -> torch.Tensor:
    """reference :273-347.  (h, w) CUDA tensors; `k` (3,3) and `c2w` (4,4) for the non-rigid method."""
    if alignment_method not in ("rigid", "non_rigid"):
        raise ValueError(f"Unsupported alignment method: {alignment_method}")
    if alignment_method == "non_rigid" and (k is None or c2w is None):
        raise ValueError("Camera intrinsics (k) and camera-to-world matrix (c2w) are required for non-rigid alignment")
    if not source_depth.is_cuda:
        raise ValueError("align_depth runs on the GPU (gen3c_b200 has no CPU path)")
    depth = _align_inv_depth_to_depth(1.0 / source_depth.float(), target_depth.float(), target_mask)
    if alignment_method == "rigid":
        return depth
    if smoothing_kernel_size != 3:
        raise NotImplementedError("the native non-rigid alignment implements the reference's 3x3 smoothing kernel")
    h, w = depth.shape
    depth = depth.contiguous()
    out = torch.empty_like(depth)
    tgt = target_depth.float().contiguous()
    m8 = (target_mask != 0).to(torch.uint8).contiguous()
    kk, cc = k.float().contiguous(), c2w.float().contiguous()
    with torch.cuda.device(depth.device):
        _lib.check(_lib.load().g3c_align_depth_nonrigid(_lib.ptr(depth), _lib.ptr(tgt), _lib.ptr(m8), _lib.ptr(kk),
                                                        _lib.ptr(cc), h, w, int(num_iters), float(lambda_arap), 0.001,
                                                        _lib.ptr(out), _lib.stream_ptr()), "g3c_align_depth_nonrigid")
    return out
