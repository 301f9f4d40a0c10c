"""ORACLE (test infrastructure only) — seeded synthetic inputs shared by oracle/make_golden.py, the
tests and bench.py.  No reference dependency; numpy / torch CPU only."""
from __future__ import annotations

import numpy as np
import torch

from .dit_oracle import DitCfg

F32 = np.float32


# --------------------------------------------------------------------------------------------------
# Path R known-answer cases (SURVEY.md §8c KAT-R1..R5, §8d config 1)
# --------------------------------------------------------------------------------------------------
def smooth_depth(h: int, w: int) -> np.ndarray:
    y, x = np.meshgrid(np.linspace(0, 1, h, dtype=F32), np.linspace(0, 1, w, dtype=F32), indexing="ij")
    return (2 + np.sin(3 * x) + 0.5 * np.cos(4 * y)).astype(F32)


def intrinsics(h: int, w: int, f: float | None = None) -> np.ndarray:
    f = f if f is not None else 200.0 * w / 256.0
    return np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]], dtype=F32)


def warp_case(name: str):
    """Returns dict(depth (b,1,h,w), image (b,3,h,w), mask|None, w2c_src (b,4,4), K (b,3,3), w2c_tgt (b,4,4))."""
    rng = np.random.RandomState({"R1": 0, "R2": 1, "R3": 2, "R4": 3, "R5": 4, "R6": 5}[name])
    eye = np.eye(4, dtype=F32)
    if name in ("R1", "R2"):  # 256x256 identity / 0.1 left pan
        h = w = 256
        depth = smooth_depth(h, w)[None, None]
        img = rng.uniform(-1, 1, (1, 3, h, w)).astype(F32)
        tgt = eye.copy()
        if name == "R2":
            tgt[0, 3] = 0.1
        return dict(depth=depth, image=img, mask=None, w2c_src=eye[None], K=intrinsics(h, w)[None], w2c_tgt=tgt[None])
    h, w = 96, 128
    K = intrinsics(h, w, 100.0)
    if name == "R3":  # chunk-of-2 coupling: two different target poses in one call
        depth = np.stack([smooth_depth(h, w), 1.5 * smooth_depth(h, w)])[:, None]
        img = rng.uniform(-1, 1, (2, 3, h, w)).astype(F32)
        # generic poses (yaw + pitch + translation): a pure axis-aligned pan puts every y coordinate
        # within 1 ulp of an integer, where floor/ceil of the reference algorithm flip with the rounding
        # order of the projection (a x2 change of that source's weight) - see DESIGN.md section 2.
        t0, t1 = look(0.03, -0.02, (-0.05, 0.01, 0.02)), look(-0.02, 0.015, (0.02, -0.01, 0.4))
        return dict(depth=depth, image=img, mask=None, w2c_src=np.stack([eye, eye]), K=np.stack([K, K]),
                    w2c_tgt=np.stack([t0, t1]))
    if name == "R4":  # behind-camera points: camera moved forward past part of the scene, with a mask
        depth = (0.3 + 1.7 * smooth_depth(h, w) / 3.5)[None, None].astype(F32)
        img = rng.uniform(-1, 1, (1, 3, h, w)).astype(F32)
        mask = (rng.uniform(0, 1, (1, 1, h, w)) > 0.2).astype(F32)
        tgt = eye.copy()
        tgt[2, 3] = -1.0
        return dict(depth=depth, image=img, mask=mask, w2c_src=eye[None], K=K[None], w2c_tgt=tgt[None])
    if name == "R5":  # integer coordinates: constant depth, shift by exactly 2 px (f*tx/z = 100*0.04/2)
        depth = np.full((1, 1, h, w), 2.0, dtype=F32)
        img = rng.uniform(-1, 1, (1, 3, h, w)).astype(F32)
        tgt = eye.copy()
        tgt[0, 3] = 0.04
        return dict(depth=depth, image=img, mask=None, w2c_src=eye[None], K=K[None], w2c_tgt=tgt[None])
    if name == "R6":  # rotation + translation, non-identity source pose
        depth = smooth_depth(h, w)[None, None]
        img = rng.uniform(-1, 1, (1, 3, h, w)).astype(F32)
        a = 0.08
        rot = np.array([[np.cos(a), 0, np.sin(a), 0.1], [0, 1, 0, -0.02], [-np.sin(a), 0, np.cos(a), 0.05],
                        [0, 0, 0, 1]], dtype=F32)
        src = eye.copy()
        src[1, 3] = 0.03
        return dict(depth=depth, image=img, mask=None, w2c_src=src[None], K=K[None], w2c_tgt=rot[None])
    raise KeyError(name)


def foreground_case():
    """R7: a near box (1.2 m) in front of a smooth background (2.9 - 4.1 m), camera yawed and
    shifted so that the box edge occludes background pixels.  boundary mask = depth-discontinuity pixels."""
    rng = np.random.RandomState(7)
    h, w = 96, 128
    depth = (2.9 + 0.35 * smooth_depth(h, w)).astype(F32)
    depth[30:70, 44:84] = 1.2 + 0.02 * smooth_depth(h, w)[30:70, 44:84]
    img = rng.uniform(-1, 1, (1, 3, h, w)).astype(F32)
    K = intrinsics(h, w, 100.0)
    eye = np.eye(4, dtype=F32)
    return dict(depth=depth[None, None], image=img, mask=None, w2c_src=eye[None], K=K[None],
                w2c_tgt=look(0.06, -0.015, (0.12, 0.01, 0.03))[None])


def look(yaw: float, pitch: float, t) -> np.ndarray:
    """world-to-camera matrix: rotation yaw (about y) then pitch (about x), translation t."""
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    m = np.eye(4)
    m[:3, :3] = rx @ ry
    m[:3, 3] = t
    return m.astype(F32)


def pan_trajectory(n: int, distance: float = 0.3) -> np.ndarray:
    """n world-to-camera matrices: a 'left' pan (translation along +x) that keeps facing the scene centre
    (yaw grows with the offset) with a slight pitch, so projected coordinates are generic (non-integer)."""
    out = []
    for i in range(n):
        a = i / max(1, n - 1)
        out.append(look(-0.12 * a * distance / 0.3 - 0.004, 0.003 + 0.01 * a, (distance * a, 0.002 * a, 0.01 * a)))
    return np.stack(out).astype(F32)


# --------------------------------------------------------------------------------------------------
# Path D cases
# --------------------------------------------------------------------------------------------------
TINY = DitCfg(model_channels=256, num_blocks=2, num_heads=2, ffn_dim=1024, context_dim=64, adaln_lora_dim=32,
              in_channels=81, out_channels=16, concat_padding_mask=True, max_frames=16, max_h=32, max_w=32,
              rope_t_ratio=2.0)
TINY_SHAPE = dict(T=2, H=16, W=16, ctx_len=128)  # L = 2*8*8 = 128 tokens


# one block at the full BASELINE width on two latent frames of the 720p grid: 2 * 44 * 80 = 7 040 tokens (55 x 128)
FULLWIDTH_1BLOCK = DitCfg(num_blocks=1)
FULLWIDTH_SHAPE = dict(T=2, H=88, W=160, ctx_len=512)


def dit_inputs(cfg: DitCfg, T: int, H: int, W: int, ctx_len: int, seed: int = 1, x_scale: float = 1.0):
    """bf16-representable fp32 tensors: x, cond_mask, cond_pose, padding_mask, ctx (cond/uncond), timestep."""
    g = torch.Generator().manual_seed(seed)

    def r(*shape, s=1.0):
        return (s * torch.randn(*shape, generator=g)).to(torch.bfloat16).float()

    x = r(16, T, H, W, s=x_scale)
    cond_mask = torch.zeros(1, T, H, W)
    cond_mask[:, 0] = 1.0
    pose = r(cfg.in_channels - 17, T, H, W, s=0.5)
    padding = torch.zeros(H, W)
    ctx_c = r(ctx_len, cfg.context_dim)
    ctx_u = r(ctx_len, cfg.context_dim)
    gt = r(16, T, H, W, s=0.5)
    timestep = float(torch.tensor(0.734375))  # exactly representable in bf16
    return dict(x=x, cond_mask=cond_mask, pose=pose, padding=padding, ctx_c=ctx_c, ctx_u=ctx_u, gt=gt,
                timestep=timestep)


# --------------------------------------------------------------------------------------------------
# a tiny TorchScript tokenizer checkpoint (encoder.jit / decoder.jit / mean_std.pt) for the VAE-wrapper tests
# --------------------------------------------------------------------------------------------------
class _TinyEnc(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.w = torch.nn.Parameter(torch.randn(16, 3, generator=g))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # [B,3,1+8k,H,W] -> [B,16,1+k,H/8,W/8]
        b0, c0, t0, h0, w0 = x.shape
        xs = x.reshape(b0, c0, t0, h0 // 8, 8, w0 // 8, 8).mean(6).mean(4)
        first, rest = xs[:, :, :1], xs[:, :, 1:]
        b, c, t, h, w = rest.shape
        rest = rest.reshape(b, c, t // 8, 8, h, w).mean(3)
        return torch.einsum("oc,bcthw->bothw", self.w, torch.cat([first, rest], 2))


class _TinyDec(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(4)
        self.w = torch.nn.Parameter(torch.randn(3, 16, generator=g) * 0.3)

    def forward(self, z: torch.Tensor) -> torch.Tensor:  # [B,16,1+k,h,w] -> [B,3,1+8k,8h,8w]
        rgb = torch.einsum("oc,bcthw->bothw", self.w, z)
        t = torch.cat([rgb[:, :, :1], rgb[:, :, 1:].repeat_interleave(8, dim=2)], 2)
        return torch.nn.functional.interpolate(t, scale_factor=(1.0, 8.0, 8.0), mode="nearest")


def write_tiny_tokenizer(vae_dir: str) -> None:
    """encoder.jit, decoder.jit, mean_std.pt in the layout of checkpoints/Cosmos-Tokenize1-CV8x8x8-720p."""
    import os

    os.makedirs(vae_dir, exist_ok=True)
    torch.jit.script(_TinyEnc()).save(os.path.join(vae_dir, "encoder.jit"))
    torch.jit.script(_TinyDec()).save(os.path.join(vae_dir, "decoder.jit"))
    g = torch.Generator().manual_seed(5)
    mean = torch.randn(16 * 16, generator=g) * 0.2
    std = 0.5 + torch.rand(16 * 16, generator=g)
    torch.save((mean, std), os.path.join(vae_dir, "mean_std.pt"))


def tiny_tokenizer_video() -> torch.Tensor:
    g = torch.Generator().manual_seed(6)
    return (torch.rand(1, 3, 34, 16, 32, generator=g) * 2 - 1)  # two chunks of 17 frames
