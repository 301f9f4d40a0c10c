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


def project_points(world_points, w2c, intrinsic, return_cam_points=False):
    """:462-486.  world_points (b,h,w,3) -> (b,h,w,3) = K (w2c [p;1])[:3]; optionally also the camera-space points."""
    b, h, w, _ = world_points.shape
    homo = np.concatenate([world_points.astype(F32), np.ones((b, h, w, 1), dtype=F32)], axis=3)[..., None]
    cam = np.matmul(w2c.astype(F32)[:, None, None], homo)[:, :, :, :3]
    proj = np.matmul(intrinsic.astype(F32)[:, None, None], cam)
    if return_cam_points:
        return proj[..., 0].astype(F32), cam[..., 0].astype(F32)
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
                 render_depth=False, foreground_masking=False, boundary_mask=None):
    """:171-336, the depth1=None / world_points1 branch (:219-224, :244-250, :281-284), without
    normal filtering; with the foreground-masking occlusion pass (:285-335) when asked.
    Returns (warped, mask2, depth2|None, flow12)."""
    frame1 = frame1.astype(F32)
    b, c, h, w = frame1.shape
    if mask1 is None:
        mask1 = np.ones((b, 1, h, w), dtype=F32)
    tp, cam_points_target = project_points(world_points1, transformation2, intrinsic2, return_cam_points=True)
    z = tp[:, :, :, 2][:, None]
    mask1 = mask1.astype(F32) * (z > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        coords = tp[:, :, :, :2] / (tp[:, :, :, 2:3] + F32(1e-7))
    coords = np.moveaxis(coords, 3, 1).astype(F32)
    flow12 = coords - create_grid(b, h, w)
    warped, mask2 = bilinear_splatting(frame1, mask1, z, flow12, is_image=is_image)
    depth2 = None
    if render_depth or foreground_masking:
        depth2 = bilinear_splatting(z, mask1, z, flow12, is_image=False)[0][:, 0]
    if foreground_masking:
        assert boundary_mask is not None
        for bi in range(b):
            closer = foreground_occlusion(cam_points_target[bi], boundary_mask[bi].astype(bool), intrinsic2[bi], depth2[bi])
            if closer is None:
                continue
            keep = (~closer).astype(F32)
            mask2[bi, 0] = mask2[bi, 0] * keep
            warped[bi] = (warped[bi] + F32(1)) * keep[None] - F32(1)
            depth2[bi] = depth2[bi] * keep
    return warped, mask2, depth2, flow12


# ------------------------------------------------------------------------------------------------------------------
# foreground-masking occlusion pass
# ------------------------------------------------------------------------------------------------------------------
def _interp_axis_bilinear(n_in: int, n_out: int):
    """Source indices / weights of F.interpolate(mode="bilinear", align_corners=False) along one axis."""
    scale = F32(n_in) / F32(n_out)
    src = np.maximum((np.arange(n_out, dtype=F32) + F32(0.5)) * scale - F32(0.5), F32(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = (src - i0.astype(F32)).astype(F32)
    return i0, i1, lam


def resize_bilinear(x: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """x (..., H, W) -> (..., new_h, new_w), torch's bilinear, align_corners=False (no antialias)."""
    y0, y1, ly = _interp_axis_bilinear(x.shape[-2], new_h)
    x0, x1, lx = _interp_axis_bilinear(x.shape[-1], new_w)
    x = x.astype(F32)
    top = x[..., y0, :][..., :, x0] * (F32(1) - lx) + x[..., y0, :][..., :, x1] * lx
    bot = x[..., y1, :][..., :, x0] * (F32(1) - lx) + x[..., y1, :][..., :, x1] * lx
    return (top * (F32(1) - ly)[:, None] + bot * ly[:, None]).astype(F32)


def resize_nearest(x: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """torch's mode="nearest": source index floor(dst * in / out)."""
    yi = np.minimum(np.floor(np.arange(new_h, dtype=F32) * (F32(x.shape[-2]) / F32(new_h))).astype(np.int64), x.shape[-2] - 1)
    xi = np.minimum(np.floor(np.arange(new_w, dtype=F32) * (F32(x.shape[-1]) / F32(new_w))).astype(np.int64), x.shape[-1] - 1)
    return x[..., yi, :][..., :, xi]


def points_to_mesh(points: np.ndarray, mask: np.ndarray, resolution=None):
    """forward_warp_utils_pytorch.py:49-132.  points (H,W,3), mask (H,W) bool -> (vertices (H'*W',3), faces (M,3) int64).
    The reference additionally drops unused vertices and renumbers the faces; that relabelling does not change any
    triangle, so the vertex grid is kept whole here."""
    if resolution is not None:
        nh, nw = resolution
        points = np.moveaxis(resize_bilinear(np.moveaxis(points.astype(F32), 2, 0), nh, nw), 0, 2)
        mask = resize_nearest(mask.astype(F32), nh, nw) != 0
    H, W = mask.shape
    idx = np.arange(H * W).reshape(H, W)
    valid = mask[:-1, :-1] | mask[:-1, 1:] | mask[1:, :-1] | mask[1:, 1:]
    vh, vw = np.nonzero(valid)
    tl, tr, bl, br = idx[vh, vw], idx[vh, vw + 1], idx[vh + 1, vw], idx[vh + 1, vw + 1]
    faces = np.concatenate([np.stack([tl, tr, bl], 1), np.stack([tr, br, bl], 1)], 0)
    return points.reshape(-1, 3).astype(F32), faces


def get_camera_rays(h: int, w: int, intrinsic: np.ndarray) -> np.ndarray:
    """:151-168.  intrinsic (3,3) -> unit rays (h,w,3) through the pixel centres' integer coordinates."""
    kinv = inverse_with_conversion(intrinsic[None])[0]
    xs, ys = np.meshgrid(np.arange(w, dtype=F32), np.arange(h, dtype=F32))
    pos = np.stack([xs, ys, np.ones_like(xs)], axis=-1)[..., None]  # (h,w,3,1)
    un = np.matmul(kinv[None, None], pos)[..., 0].astype(F32)
    nrm = np.linalg.norm(un, axis=-1, keepdims=True).astype(F32)
    nrm[nrm == 0] = 1
    return (un / nrm).astype(F32)


def ray_triangle_depth(origins: np.ndarray, dirs: np.ndarray, vertices: np.ndarray, faces: np.ndarray,
                       eps: float = 1e-8, tri_chunk: int = 512) -> np.ndarray:
    """Restates the NVIDIA Warp kernel ray_triangle_intersection_warp.py:23-105 (Moeller-Trumbore, nearest hit with
    t > eps, 0 where nothing is hit) in float32.  PARITY UNPINNED for this function: the reference implementation is a
    Warp/CUDA kernel that cannot run in the build container; everything around it is pinned (make_golden.py)."""
    o, d = origins.reshape(-1, 3).astype(F32), dirs.reshape(-1, 3).astype(F32)
    best = np.full(o.shape[0], F32(1e10), dtype=F32)
    eps = F32(eps)
    for t0 in range(0, faces.shape[0], tri_chunk):
        f = faces[t0:t0 + tri_chunk]
        v0, v1, v2 = vertices[f[:, 0]], vertices[f[:, 1]], vertices[f[:, 2]]
        e1, e2 = (v1 - v0).astype(F32), (v2 - v0).astype(F32)                   # (T,3)
        h = np.cross(d[:, None, :], e2[None, :, :]).astype(F32)                  # (R,T,3)
        a = np.einsum("tk,rtk->rt", e1, h).astype(F32)
        ok = np.abs(a) >= eps
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            fi = (F32(1) / a).astype(F32)
            s = (o[:, None, :] - v0[None, :, :]).astype(F32)
            u = (fi * np.einsum("rtk,rtk->rt", s, h)).astype(F32)
            ok &= ~((u < 0) | (u > 1))
            q = np.cross(s, e1[None, :, :]).astype(F32)
            v = (fi * np.einsum("rk,rtk->rt", d, q)).astype(F32)
            ok &= ~((v < 0) | ((u + v) > 1))
            t = (fi * np.einsum("tk,rtk->rt", e2, q)).astype(F32)
        ok &= t > eps
        t = np.where(ok, t, F32(1e10))
        best = np.minimum(best, t.min(axis=1))
    return np.where(best < F32(1e10), best, F32(0)).astype(F32)


def foreground_occlusion(cam_points_target, boundary_mask, intrinsic2, warped_depth2, mesh_downsample_factor=4):
    """:285-329 for one batch item: mesh of the 1/4-resolution target-camera points around boundary pixels, nearest
    ray/mesh hit per target pixel, `closer` = mesh in front of the splatted depth by more than 0.02.  None when the
    mesh is empty (the reference `continue`s)."""
    h, w = boundary_mask.shape
    vertices, faces = points_to_mesh(cam_points_target, boundary_mask, (h // mesh_downsample_factor, w // mesh_downsample_factor))
    if faces.shape[0] == 0:
        return None
    rays = get_camera_rays(h, w, intrinsic2)
    t = ray_triangle_depth(np.zeros_like(rays), rays, vertices, faces).reshape(h, w)
    mesh_z = resize_bilinear((t * rays[:, :, 2]).astype(F32)[None], h, w)[0]  # same size: identity resample
    return ((mesh_z + F32(0.02)) < warped_depth2) & (mesh_z > 0)


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


# --------------------------------------------------------------------------------------------------
# depth alignment of Cache3D_Buffer.update_cache
# --------------------------------------------------------------------------------------------------
def _quantile(x: np.ndarray, q) -> np.ndarray:
    """torch.quantile(x, q) with the default linear interpolation, float32."""
    return np.quantile(x.astype(F32), np.asarray(q, dtype=F32), method="linear").astype(F32)


def align_inv_depth_to_depth(source_inv_depth, target_depth, target_mask=None):
    """camera_utils.py:225-270 — affine fit (scale, bias) of the source inverse depth to the target inverse depth on the
    10 %..90 % quantile core of both, solved by least squares; returns the aligned DEPTH (h, w)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        target_inv = (F32(1.0) / target_depth.astype(F32)).astype(F32)
    src_mask = source_inv_depth > 0
    tdm = target_depth > 0
    tmask = tdm if target_mask is None else np.logical_and(target_mask > 0, tdm)
    s_lo, s_hi = _quantile(source_inv_depth[src_mask], [0.1, 0.9])
    t_lo, t_hi = _quantile(target_inv[tmask], [0.1, 0.9])
    src_mask = (source_inv_depth > s_lo) & (source_inv_depth < s_hi)
    tmask = (target_inv > t_lo) & (target_inv < t_hi)
    m = src_mask & tmask
    a = np.stack([source_inv_depth[m].astype(np.float64), np.ones(int(m.sum()))], axis=1)
    sol = np.linalg.lstsq(a, target_inv[m].astype(np.float64), rcond=None)[0]
    scale, bias = F32(sol[0]), F32(sol[1])
    with np.errstate(divide="ignore"):
        return (F32(1.0) / (source_inv_depth.astype(F32) * scale + bias)).astype(F32)


def align_depth(source_depth, target_depth, target_mask, k=None, c2w=None, alignment_method="rigid", num_iters=100,
                lambda_arap=0.1, smoothing_kernel_size=3, lr=0.001):
    """camera_utils.py:273-347.  rigid: the affine inverse-depth fit.  non_rigid: a per-pixel scale map, 100 Adam steps
    (lr 1e-3, betas .9/.999, eps 1e-8) on  mean|unproject(src*sc) - unproject(tgt)| over the masked pixels
    + lambda * mean|box3(sc) - sc|.  The reference differentiates with autograd; here the gradient is written out:
      d data / d sc_p = sign(e_p) * d_p * (|v_p|_1) / (3 n),  e_p = d_p sc_p - t_p,  v_p = R K^-1 (x, y, 1)
      d arap / d sc   = (box3(g) - g) / (H W),                 g = sign(box3(sc) - sc)      (zero padding)
    (R = rotation of inverse(c2w): the reference passes c2w where unproject_points expects w2c, :298-305,:316-322)."""
    src = source_depth.astype(F32)
    with np.errstate(divide="ignore"):
        depth = align_inv_depth_to_depth((F32(1.0) / src).astype(F32), target_depth.astype(F32), target_mask)
    if alignment_method == "rigid":
        return depth
    assert alignment_method == "non_rigid" and k is not None and c2w is not None and smoothing_kernel_size == 3
    h, w = depth.shape
    mask = target_mask.astype(bool)
    n = int(mask.sum())
    kinv = inverse_with_conversion(k)
    rot = inverse_with_conversion(c2w)[:3, :3]
    ys, xs = np.meshgrid(np.arange(h, dtype=F32), np.arange(w, dtype=F32), indexing="ij")
    rays = np.stack([xs, ys, np.ones_like(xs)], -1) @ kinv.T        # (h, w, 3)
    v1 = np.abs(rays @ rot.T).sum(-1).astype(F32)                    # |R r|_1
    coef = np.where(mask, depth * v1 / F32(3 * max(n, 1)), F32(0)).astype(F32)
    tgt = target_depth.astype(F32)

    def box3(a):
        p = np.pad(a, 1)
        s = np.zeros_like(a)
        for dy in range(3):
            for dx in range(3):
                s = s + p[dy:dy + h, dx:dx + w] * F32(1.0 / 9.0)
        return s.astype(F32)

    sc = np.ones((h, w), F32)
    m1 = np.zeros((h, w), F32)
    m2 = np.zeros((h, w), F32)
    b1, b2, eps = 0.9, 0.999, 1e-8
    for it in range(1, num_iters + 1):
        e = depth * sc - tgt
        g = np.sign(box3(sc) - sc).astype(F32)
        grad = coef * np.sign(e) + F32(lambda_arap / (h * w)) * (box3(g) - g)
        m1 = (b1 * m1 + (1 - b1) * grad).astype(F32)
        m2 = (b2 * m2 + (1 - b2) * grad * grad).astype(F32)
        denom = np.sqrt(m2) / F32(np.sqrt(1 - b2 ** it)) + F32(eps)
        sc = (sc - F32(lr / (1 - b1 ** it)) * m1 / denom).astype(F32)
    return (depth * sc).astype(F32)
