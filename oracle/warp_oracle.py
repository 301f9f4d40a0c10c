"""ORACLE (test infrastructure only) — numpy restatement of GEN3C's 3D-cache render arithmetic.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product path (``gen3c_b200``) never does.

Each function restates, in float32 numpy, one function of the reference
``cosmos_predict1/diffusion/inference/forward_warp_utils_pytorch.py`` (cited per function) and
``cache_3d.py``.  Pinned against the reference's own code executed in this container
(``oracle/make_golden.py`` -> ``tests/golden/warp_*.npz``): the reference ships no golden vectors
for this path, so the fixtures are minted from the reference itself.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def create_grid(b: int, h: int, w: int) -> np.ndarray:
    """forward_warp_utils_pytorch.py:697-703 -> (b, 2, h, w) of (x, y)."""
    x = np.broadcast_to(np.arange(w, dtype=F32).reshape(1, 1, 1, w), (b, 1, h, w))
    y = np.broadcast_to(np.arange(h, dtype=F32).reshape(1, 1, h, 1), (b, 1, h, w))
    return np.concatenate([x, y], axis=1)


def inverse_with_conversion(m: np.ndarray) -> np.ndarray:
    """:147-148 — torch.linalg.inv in float32."""
    return np.linalg.inv(m.astype(F32)).astype(F32)


def unproject_points(depth, w2c, intrinsic, is_depth=True, mask=None) -> np.ndarray:
    """:410-460.  depth (b,1,h,w), w2c (b,4,4), intrinsic (b,3,3) -> (b,h,w,3)."""
    depth = depth.astype(F32)
    b, _, h, w = depth.shape
    if mask is None:
        mask = depth > 0
    if mask.ndim == depth.ndim and mask.shape[1] == 1:
        mask = mask[:, 0]
    out = np.zeros((b, h, w, 3), dtype=F32)
    bi, yi, xi = np.nonzero(mask)
    if bi.size == 0:
        return out
    kinv = inverse_with_conversion(intrinsic)
    pos = np.stack([xi.astype(F32), yi.astype(F32), np.ones_like(xi, dtype=F32)], axis=1)[..., None]
    unnorm = np.matmul(kinv[bi], pos)  # (N,3,1)
    dv = depth[bi, 0, yi, xi].reshape(-1, 1, 1)
    if is_depth:
        cam = dv * unnorm
    else:
        nrm = np.linalg.norm(unnorm, axis=1, keepdims=True).astype(F32)
        cam = dv * (unnorm / (nrm + F32(1e-8)))
    homo = np.concatenate([cam, np.ones((cam.shape[0], 1, 1), dtype=F32)], axis=1)
    c2w = inverse_with_conversion(w2c)
    world = np.matmul(c2w[bi], homo)
    out[bi, yi, xi, :] = world[:, :3, 0]
    return out


def project_points(world_points, w2c, intrinsic):
    """:462-486.  world_points (b,h,w,3) -> (b,h,w,3) = K (w2c [p;1])[:3]."""
    b, h, w, _ = world_points.shape
    homo = np.concatenate([world_points.astype(F32), np.ones((b, h, w, 1), dtype=F32)], axis=3)[..., None]
    cam = np.matmul(w2c.astype(F32)[:, None, None], homo)[:, :, :, :3]
    proj = np.matmul(intrinsic.astype(F32)[:, None, None], cam)
    return proj[..., 0].astype(F32)


def splat_indices(flow):
    """The integer part of bilinear_splatting (:605-621): floor/ceil of (flow+grid)+1 taken BEFORE
    clamping, then clamped to x in [0,w+1], y in [0,h+1].  Returns (pos_clamped, floor, ceil) with
    floor/ceil int64 (b,2,h,w)."""
    b, _, h, w = flow.shape
    grid = create_grid(b, h, w)
    pos = (flow.astype(F32) + grid) + F32(1)
    with np.errstate(invalid="ignore"):
        fl = np.floor(pos)
        ce = np.ceil(pos)
        # torch .long() of NaN/inf on the GPU saturates; keep the finite case exact and send
        # non-finite coordinates to the (cropped) border like the CUDA path does.
        fl = np.nan_to_num(fl, nan=0.0, posinf=1e9, neginf=-1e9).astype(np.int64)
        ce = np.nan_to_num(ce, nan=0.0, posinf=1e9, neginf=-1e9).astype(np.int64)
    lim = np.array([w + 1, h + 1]).reshape(1, 2, 1, 1)
    posc = np.clip(np.nan_to_num(pos, nan=0.0), 0, lim.astype(F32)).astype(F32)
    fl = np.clip(fl, 0, lim)
    ce = np.clip(ce, 0, lim)
    return posc, fl, ce


def bilinear_splatting(frame1, mask1, depth1, flow12, is_image=False, depth_weight_scale=50):
    """:576-695 (flow12_mask=None, n_views=1).  Returns (warped (b,c,h,w), mask (b,1,h,w))."""
    frame1 = frame1.astype(F32)
    b, c, h, w = frame1.shape
    if mask1 is None:
        mask1 = np.ones((b, 1, h, w), dtype=F32)
    mask1 = mask1.astype(F32)
    depth1 = depth1.astype(F32)
    pos, fl, ce = splat_indices(flow12)
    flf, cef = fl.astype(F32), ce.astype(F32)
    one = F32(1)
    w_nw = (one - (pos[:, 1:2] - flf[:, 1:2])) * (one - (pos[:, 0:1] - flf[:, 0:1]))
    w_sw = (one - (cef[:, 1:2] - pos[:, 1:2])) * (one - (pos[:, 0:1] - flf[:, 0:1]))
    w_ne = (one - (pos[:, 1:2] - flf[:, 1:2])) * (one - (cef[:, 0:1] - pos[:, 0:1]))
    w_se = (one - (cef[:, 1:2] - pos[:, 1:2])) * (one - (cef[:, 0:1] - pos[:, 0:1]))
    logd = np.log1p(np.maximum(depth1, F32(0))).astype(F32)
    expo = logd / (logd.max() + F32(1e-7)) * F32(depth_weight_scale)
    expo = np.minimum(expo, F32(80.0))
    dw = np.exp(expo).astype(F32) + F32(1e-7)
    acc = np.zeros((b, h + 2, w + 2, c), dtype=F32)
    wsum = np.zeros((b, h + 2, w + 2, 1), dtype=F32)
    frame_cl = np.moveaxis(frame1, 1, 3)
    bidx = np.arange(b)[:, None, None]
    for wt, yy, xx in ((w_nw, fl[:, 1], fl[:, 0]), (w_sw, ce[:, 1], fl[:, 0]),
                       (w_ne, fl[:, 1], ce[:, 0]), (w_se, ce[:, 1], ce[:, 0])):
        wgt = np.moveaxis((wt * mask1 / dw).astype(F32), 1, 3)  # (b,h,w,1)
        np.add.at(acc, (bidx, yy, xx), (frame_cl * wgt).astype(F32))
        np.add.at(wsum, (bidx, yy, xx), wgt)
    acc = np.moveaxis(acc, 3, 1)[:, :, 1:-1, 1:-1]
    ws = np.moveaxis(wsum, 3, 1)[:, :, 1:-1, 1:-1]
    ws = np.nan_to_num(ws, nan=1000.0)
    hit = ws > 0
    zero = F32(-1 if is_image else 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(hit, acc / ws, zero).astype(F32)
    if is_image:
        out = np.clip(out, -1, 1)
    return out, hit.astype(F32)


def forward_warp(frame1, mask1, world_points1, transformation2, intrinsic2, is_image=True,
                 render_depth=False):
    """:171-336, the depth1=None / world_points1 branch (:219-224, :244-250, :281-284), without
    normal filtering or foreground masking.  Returns (warped, mask2, depth2|None, flow12)."""
    frame1 = frame1.astype(F32)
    b, c, h, w = frame1.shape
    if mask1 is None:
        mask1 = np.ones((b, 1, h, w), dtype=F32)
    tp = project_points(world_points1, transformation2, intrinsic2)  # (b,h,w,3)
    z = tp[:, :, :, 2][:, None]
    mask1 = mask1.astype(F32) * (z > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        coords = tp[:, :, :, :2] / (tp[:, :, :, 2:3] + F32(1e-7))
    coords = np.moveaxis(coords, 3, 1).astype(F32)
    flow12 = coords - create_grid(b, h, w)
    warped, mask2 = bilinear_splatting(frame1, mask1, z, flow12, is_image=is_image)
    depth2 = None
    if render_depth:
        depth2 = bilinear_splatting(z, mask1, z, flow12, is_image=False)[0][:, 0]
    return warped, mask2, depth2, flow12


def reliable_depth_mask_range_batch(depth, window_size=5, ratio_thresh=0.05, eps=1e-6):
    """:338-353 (max/min pools ignore padding; avg pool counts zero padding)."""
    assert window_size % 2 == 1, "Window size must be odd."
    d = depth.astype(F32)
    if d.ndim == 3:
        d = d[:, None]
    b, _, h, w = d.shape
    r = window_size // 2
    pad_hi = np.pad(d, ((0, 0), (0, 0), (r, r), (r, r)), constant_values=-np.inf)
    pad_lo = np.pad(d, ((0, 0), (0, 0), (r, r), (r, r)), constant_values=np.inf)
    pad_0 = np.pad(d, ((0, 0), (0, 0), (r, r), (r, r)), constant_values=0)
    mx = np.full_like(d, -np.inf)
    mn = np.full_like(d, np.inf)
    sm = np.zeros_like(d)
    for dy in range(window_size):
        for dx in range(window_size):
            mx = np.maximum(mx, pad_hi[:, :, dy:dy + h, dx:dx + w])
            mn = np.minimum(mn, pad_lo[:, :, dy:dy + h, dx:dx + w])
            sm = sm + pad_0[:, :, dy:dy + h, dx:dx + w]
    mean = sm / F32(window_size * window_size)
    ratio = (mx - mn) / (mean + F32(eps))
    return (ratio < ratio_thresh) & (d > 0)


def render_cache(points, images, masks, w2cs, Ks, render_depth=False, chunk=2):
    """cache_3d.py:151-236 for V=1: points (B,Fs,N,H,W,3), images (B,Fs,N,3,H,W), masks
    (B,Fs,N,1,H,W)|None with Fs in {1, F}; w2cs (B,F,4,4), Ks (B,F,3,3).
    Items are flattened (B F N) and warped in chunks of 2 that share one log-depth max.
    Returns pixels (B,F,N,3,H,W) [or depth (B,F,N,H,W)], masks (B,F,N,1,H,W)."""
    B, Fs, N, H, W, _ = points.shape
    F = w2cs.shape[1]

    def expand(a):
        return np.broadcast_to(a, (B, F) + a.shape[2:]).reshape((B * F * N,) + a.shape[3:])

    pts = expand(points)
    img = expand(images)
    msk = expand(masks) if masks is not None else None
    w2 = np.broadcast_to(w2cs[:, :, None], (B, F, N, 4, 4)).reshape(-1, 4, 4)
    kk = np.broadcast_to(Ks[:, :, None], (B, F, N, 3, 3)).reshape(-1, 3, 3)
    outs, mouts, douts = [], [], []
    for i in range(0, pts.shape[0], chunk):
        s = slice(i, i + chunk)
        wi, mi, di, _ = forward_warp(img[s], None if msk is None else msk[s], pts[s], w2[s], kk[s],
                                     render_depth=render_depth)
        outs.append(wi)
        mouts.append(mi)
        if render_depth:
            douts.append(di)
    pix = np.concatenate(outs, 0).reshape(B, F, N, 3, H, W)
    mk = np.concatenate(mouts, 0).reshape(B, F, N, 1, H, W)
    if render_depth:
        return np.concatenate(douts, 0).reshape(B, F, N, H, W), mk
    return pix, mk
