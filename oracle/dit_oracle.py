"""ORACLE (test infrastructure only) — fp32 torch-CPU restatement of the GEN3C DiT forward and of the
denoise-step loop body.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module; the product path never does.

Restates (file:line into /root/reference/cosmos_predict1/diffusion):
  VideoExtendGeneralDIT.forward   networks/general_dit_video_conditioned.py:58-217
  GeneralDIT.forward & helpers    networks/general_dit.py:272-358,439-522
  PatchEmbed / FinalLayer / Timesteps / TimestepEmbedding / DITBuildingBlock   module/blocks.py
  Attention.cal_qkv / cal_attn, GPT2FeedForward, normalize                     module/attention.py
  VideoRopePosition3DEmb / LearnablePosEmbAxis                                 module/position_embedding.py
  loop body of generate_samples_from_batch                                     model/model_v2w.py:130-149
Third-party arithmetic absent from /root/reference, restated from its published behaviour:
  transformer-engine 1.12.0 RMSNorm / fused RoPE (rotate-half) / DotProductAttention (softmax(QK^T/sqrt d)V),
  diffusers 0.32.2 EDMEulerScheduler (Karras sigmas, c_skip/c_out/c_in, Euler step).
Pinned against the reference's own graph code executed on CPU (oracle/make_golden.py ->
tests/golden/dit_*.npz).  The TE / diffusers stand-ins are themselves restatements: PARITY UNPINNED for
those two libraries until validated on a machine that has them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class DitCfg:
    model_channels: int = 4096
    num_blocks: int = 28
    num_heads: int = 32
    ffn_dim: int = 16384
    context_dim: int = 1024
    adaln_lora_dim: int = 256
    in_channels: int = 81
    out_channels: int = 16
    concat_padding_mask: bool = True
    max_frames: int = 128
    max_h: int = 120
    max_w: int = 120
    rope_h_ratio: float = 1.0
    rope_w_ratio: float = 1.0
    rope_t_ratio: float = 2.0
    base_fps: int = 24


def state_dict_shapes(cfg: DitCfg) -> dict[str, tuple[int, ...]]:
    """Reference state-dict layout (probed from the reference class)."""
    D, R, Fd, C = cfg.model_channels, cfg.adaln_lora_dim, cfg.ffn_dim, cfg.context_dim
    kin = (cfg.in_channels + (1 if cfg.concat_padding_mask else 0)) * 4
    sd = {
        "x_embedder.proj.1.weight": (D, kin),
        "pos_embedder.seq": (max(cfg.max_h, cfg.max_w, cfg.max_frames),),
        "extra_pos_embedder.pos_emb_h": (cfg.max_h, D),
        "extra_pos_embedder.pos_emb_w": (cfg.max_w, D),
        "extra_pos_embedder.pos_emb_t": (cfg.max_frames, D),
        "t_embedder.1.linear_1.weight": (D, D),
        "t_embedder.1.linear_2.weight": (3 * D, D),
        "final_layer.linear.weight": (cfg.out_channels * 4, D),
        "final_layer.adaLN_modulation.1.weight": (R, D),
        "final_layer.adaLN_modulation.2.weight": (2 * D, R),
        "affline_norm.weight": (D,),
    }
    for i in range(cfg.num_blocks):
        for j in range(3):
            p = f"blocks.block{i}.blocks.{j}."
            sd[p + "adaLN_modulation.1.weight"] = (R, D)
            sd[p + "adaLN_modulation.2.weight"] = (3 * D, R)
            if j < 2:
                k_in = D if j == 0 else C
                sd[p + "block.attn.to_q.0.weight"] = (D, D)
                sd[p + "block.attn.to_q.1.weight"] = (128,)
                sd[p + "block.attn.to_k.0.weight"] = (D, k_in)
                sd[p + "block.attn.to_k.1.weight"] = (128,)
                sd[p + "block.attn.to_v.0.weight"] = (D, k_in)
                sd[p + "block.attn.to_out.0.weight"] = (D, D)
            else:
                sd[p + "block.layer1.weight"] = (Fd, D)
                sd[p + "block.layer2.weight"] = (D, Fd)
    return sd


def random_state_dict(cfg: DitCfg, seed: int = 0, std: float = 0.02, dtype=torch.float32) -> dict[str, torch.Tensor]:
    """N(0, std^2) weights in the real layout, values rounded to bf16 so every implementation sees the
    same numbers; RMSNorm / affine gammas ~ 1 + small noise; adaLN heads random (the reference
    zero-initialises them, general_dit.py:198-203, which would make every block the identity)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(cfg).items():
        if k == "pos_embedder.seq":
            sd[k] = torch.arange(shp[0], dtype=torch.float32)
            continue
        if k.endswith(".1.weight") and len(shp) == 1 or k == "affline_norm.weight":
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = std * torch.randn(shp, generator=g)
            if "adaLN_modulation.2" in k or "linear_2" in k:
                t = t * 2.0
        sd[k] = t.to(torch.bfloat16).to(dtype)
    return sd


def random_state_dict_on(cfg: DitCfg, device, seed: int = 0, std: float = 0.02, dtype=torch.float32):
    """Same distribution as `random_state_dict`, drawn with the generator of `device` (a 7B-parameter set takes
    seconds on the GPU instead of minutes on the host).  Values are bf16-representable; the numbers differ from the
    CPU generator's, so fixtures minted on the CPU must keep using `random_state_dict`."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, shp in state_dict_shapes(cfg).items():
        if k == "pos_embedder.seq":
            sd[k] = torch.arange(shp[0], dtype=torch.float32, device=device)
            continue
        if k.endswith(".1.weight") and len(shp) == 1 or k == "affline_norm.weight":
            t = 1.0 + 0.1 * torch.randn(shp, generator=g, device=device)
        else:
            t = std * torch.randn(shp, generator=g, device=device)
            if "adaLN_modulation.2" in k or "linear_2" in k:
                t = t * 2.0
        sd[k] = t.to(torch.bfloat16).to(dtype)
    return sd


# --------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------
def timestep_sinusoid(t: torch.Tensor, D: int) -> torch.Tensor:
    """blocks.py:38-51: [cos(t*e) | sin(t*e)], e_i = exp(-ln(1e4) * i / half)."""
    half = D // 2
    e = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t.float()[:, None] * e[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    xf = x.float()
    return xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * w.float()


def rope_angles(cfg: DitCfg, T: int, Hp: int, Wp: int, fps: float, t0: int = 0, device=None) -> torch.Tensor:
    """position_embedding.py:106-187 -> [T*Hp*Wp, 128] fp32 angles (t | h | w) repeated twice."""
    dim = 128
    dim_h = dim // 6 * 2
    dim_w = dim_h
    dim_t = dim - 2 * dim_h
    sr = torch.arange(0, dim_h, 2)[: dim_h // 2].float() / dim_h
    tr = torch.arange(0, dim_t, 2)[: dim_t // 2].float() / dim_t
    h_theta = 10000.0 * cfg.rope_h_ratio ** (dim_h / (dim_h - 2))
    w_theta = 10000.0 * cfg.rope_w_ratio ** (dim_w / (dim_w - 2))
    t_theta = 10000.0 * cfg.rope_t_ratio ** (dim_t / (dim_t - 2))
    hf = 1.0 / (h_theta ** sr)
    wf = 1.0 / (w_theta ** sr)
    tf = 1.0 / (t_theta ** tr)
    seq = torch.arange(max(cfg.max_h, cfg.max_w, cfg.max_frames), dtype=torch.float32)
    eh = torch.outer(seq[:Hp], hf)
    ew = torch.outer(seq[:Wp], wf)
    et = torch.outer(seq[t0:t0 + T] / fps * cfg.base_fps, tf)
    em = torch.cat([
        et[:, None, None, :].expand(T, Hp, Wp, -1),
        eh[None, :, None, :].expand(T, Hp, Wp, -1),
        ew[None, None, :, :].expand(T, Hp, Wp, -1),
    ] * 2, dim=-1)
    return em.reshape(T * Hp * Wp, dim).float().to(device)


def apply_rope(x: torch.Tensor, ang: torch.Tensor) -> torch.Tensor:
    """x [L, h, 128], ang [L, 128]: t*cos + rotate_half(t)*sin (TE sbhd, NeoX halves)."""
    cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
    d = x.shape[-1] // 2
    rot = torch.cat((-x[..., d:], x[..., :d]), dim=-1)
    return x * cos + rot * sin


def abs_pos_emb(sd, cfg: DitCfg, T: int, Hp: int, Wp: int, t0: int = 0) -> torch.Tensor:
    """position_embedding.py:220-233 + attention.py:108-124 -> [T*Hp*Wp, D]."""
    D = cfg.model_channels
    et = sd["extra_pos_embedder.pos_emb_t"].float()[t0:t0 + T]
    eh = sd["extra_pos_embedder.pos_emb_h"].float()[:Hp]
    ew = sd["extra_pos_embedder.pos_emb_w"].float()[:Wp]
    emb = et[:, None, None, :] + eh[None, :, None, :] + ew[None, None, :, :]
    norm = torch.linalg.vector_norm(emb, dim=-1, keepdim=True, dtype=torch.float32)
    norm = 1e-6 + norm * math.sqrt(norm.numel() / emb.numel())
    return (emb / norm).reshape(T * Hp * Wp, D)


def attention(q, k, v, heads: int) -> torch.Tensor:
    """q [Lq, D], k/v [Lk, D] (already normed / roped) -> [Lq, D]; softmax(QK^T / sqrt(128)) V."""
    Lq, D = q.shape
    qh = q.reshape(Lq, heads, 128).permute(1, 0, 2)
    kh = k.reshape(-1, heads, 128).permute(1, 0, 2)
    vh = v.reshape(-1, heads, 128).permute(1, 0, 2)
    if q.is_cuda and q.dtype == torch.float32:
        # fp32 on the GPU: explicit matmul / softmax in query chunks (plain fp32 FMA GEMMs; torch's fused fp32 SDPA
        # kernels may use TF32-class tensor instructions, which an oracle must not)
        Lk = kh.shape[1]
        rows = max(128, min(Lq, (1 << 31) // max(1, heads * Lk)))  # <= 8 GiB of fp32 scores per chunk
        out = torch.empty(heads, Lq, 128, dtype=q.dtype, device=q.device)
        for r0 in range(0, Lq, rows):
            sc = torch.matmul(qh[:, r0:r0 + rows], kh.transpose(1, 2)) * (128 ** -0.5)
            out[:, r0:r0 + rows] = torch.matmul(torch.softmax(sc, dim=-1), vh)
            del sc
        return out.permute(1, 0, 2).reshape(Lq, D)
    o = F.scaled_dot_product_attention(qh[None], kh[None], vh[None])[0]
    return o.permute(1, 0, 2).reshape(Lq, D)


def patchify(x: torch.Tensor) -> torch.Tensor:
    """[C, T, H, W] -> [T*Hp*Wp, C*4], "c (t r)(h m)(w n) -> t h w (c r m n)", r=1, m=n=2."""
    C, T, H, W = x.shape
    x = x.reshape(C, T, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5)
    return x.reshape(T * (H // 2) * (W // 2), C * 4)


def unpatchify(y: torch.Tensor, T: int, Hp: int, Wp: int, C: int) -> torch.Tensor:
    """[L, (p1 p2 t C)] -> [C, T, 2Hp, 2Wp]   (general_dit.py:348-357)."""
    y = y.reshape(T, Hp, Wp, 2, 2, C).permute(5, 0, 1, 3, 2, 4)
    return y.reshape(C, T, Hp * 2, Wp * 2)


def modulation_vectors(sd, cfg: DitCfg, timestep: float, dt=torch.float32):
    D = cfg.model_channels
    w1 = sd["t_embedder.1.linear_1.weight"]
    s = timestep_sinusoid(torch.tensor([timestep], dtype=torch.float32, device=w1.device), D)[0].to(dt)
    h1 = w1.to(dt) @ s
    lora = sd["t_embedder.1.linear_2.weight"].to(dt) @ F.silu(h1)
    emb = rms_norm(s, sd["affline_norm.weight"]).to(dt)
    return s, emb, lora


def forward(sd, cfg: DitCfg, x, cond_mask, cond_pose, padding_mask, timestep: float, ctx, fps: float = 24.0,
            t0: int = 0, T_total: int | None = None, kv_gather=None, return_intermediates: bool = False,
            compute_dtype: torch.dtype = torch.float32, num_blocks: int | None = None):
    """One network forward for B=1.  x [16,T,H,W], cond_mask [1,T,H,W], cond_pose [64,T,H,W] or None
    (zeros), padding_mask [H,W] or None (zeros), ctx [M, context_dim].
    Runs on the device of `x` (weights in `sd` must live there too) in `compute_dtype`: float32 (the oracle proper;
    on CUDA the caller switches TF32 off) or bfloat16 (every tensor stored in bf16 between ops as in the
    reference's bf16 inference run; norms / softmax accumulate in fp32 inside the op, as torch / TE do).
    t0 / kv_gather emulate a context-parallel rank: positions start at latent frame t0 and
    kv_gather(k, v) returns the K/V of all ranks.  num_blocks < cfg.num_blocks stops after that many blocks
    (error-growth-by-depth measurements) and still applies the final layer."""
    dt, dev = compute_dtype, x.device
    D, heads = cfg.model_channels, cfg.num_heads
    _, T, H, W = x.shape
    Hp, Wp = H // 2, W // 2
    L = T * Hp * Wp
    npose = cfg.in_channels - 17
    nb = cfg.num_blocks if num_blocks is None else num_blocks

    def w(k):
        return sd[k].to(device=dev, dtype=dt)

    parts = [x.to(dt), cond_mask.to(dt)]
    if npose > 0:
        parts.append(cond_pose.to(dt) if cond_pose is not None else torch.zeros(npose, T, H, W, dtype=dt, device=dev))
    if cfg.concat_padding_mask:
        pm = padding_mask.to(dt) if padding_mask is not None else torch.zeros(H, W, dtype=dt, device=dev)
        parts.append(pm[None, None].expand(1, T, H, W))
    tok = patchify(torch.cat(parts, 0))
    h = tok @ w("x_embedder.proj.1.weight").T  # [L, D]
    sdv = sd if sd["extra_pos_embedder.pos_emb_t"].device == dev else {k: sd[k].to(dev) for k in (
        "extra_pos_embedder.pos_emb_t", "extra_pos_embedder.pos_emb_h", "extra_pos_embedder.pos_emb_w")}
    pos = abs_pos_emb(sdv, cfg, T, Hp, Wp, t0).to(dt)
    ang = rope_angles(cfg, T, Hp, Wp, fps, t0, device=dev)
    s, emb, lora = modulation_vectors(sd, cfg, timestep, dt)
    inter = {}

    def mod(prefix, n):
        a = w(prefix + "adaLN_modulation.1.weight") @ F.silu(emb)
        m = w(prefix + "adaLN_modulation.2.weight") @ a + lora[: n * D]
        return m.chunk(n)

    def ln(v):
        return F.layer_norm(v, (D,), eps=1e-6)

    def nrm(v, key):  # TE RMSNorm: fp32 inside, stored in the compute dtype
        return rms_norm(v, sd[key].to(dev)).to(dt)

    def rope(v):
        return apply_rope(v.float(), ang).to(dt)

    ctx = ctx.to(dt)
    for i in range(nb):
        h = h + pos
        # FA
        p = f"blocks.block{i}.blocks.0."
        shift, scale, gate = mod(p, 3)
        xn = ln(h) * (1 + scale) + shift
        q = xn @ w(p + "block.attn.to_q.0.weight").T
        k = xn @ w(p + "block.attn.to_k.0.weight").T
        v = xn @ w(p + "block.attn.to_v.0.weight").T
        q = rope(nrm(q.reshape(L, heads, 128), p + "block.attn.to_q.1.weight")).reshape(L, D)
        k = rope(nrm(k.reshape(L, heads, 128), p + "block.attn.to_k.1.weight")).reshape(L, D)
        if kv_gather is not None:
            k, v = kv_gather(i, k, v)
        o = attention(q, k, v, heads)
        h = h + gate * (o @ w(p + "block.attn.to_out.0.weight").T)
        del q, k, v, o
        if return_intermediates and i == 0:
            inter["fa0_x"] = h.clone()
        # CA
        p = f"blocks.block{i}.blocks.1."
        shift, scale, gate = mod(p, 3)
        xn = ln(h) * (1 + scale) + shift
        q = xn @ w(p + "block.attn.to_q.0.weight").T
        kc = ctx @ w(p + "block.attn.to_k.0.weight").T
        vc = ctx @ w(p + "block.attn.to_v.0.weight").T
        q = nrm(q.reshape(L, heads, 128), p + "block.attn.to_q.1.weight").reshape(L, D)
        kc = nrm(kc.reshape(-1, heads, 128), p + "block.attn.to_k.1.weight").reshape(-1, D)
        o = attention(q, kc, vc, heads)
        h = h + gate * (o @ w(p + "block.attn.to_out.0.weight").T)
        del q, o
        # MLP
        p = f"blocks.block{i}.blocks.2."
        shift, scale, gate = mod(p, 3)
        xn = ln(h) * (1 + scale) + shift
        hid = F.gelu(xn @ w(p + "block.layer1.weight").T)
        h = h + gate * (hid @ w(p + "block.layer2.weight").T)
        del hid, xn
    shift, scale = mod("final_layer.", 2)
    y = (ln(h) * (1 + scale) + shift) @ w("final_layer.linear.weight").T
    out = unpatchify(y, T, Hp, Wp, cfg.out_channels)
    if return_intermediates:
        return out, inter
    return out


# --------------------------------------------------------------------------------------------------
# sampler (EDM Euler; diffusers 0.32.2 semantics restated — parity unpinned, see module docstring)
# --------------------------------------------------------------------------------------------------
def karras_sigmas(num_steps: int, sigma_max: float = 80.0, sigma_min: float = 0.0002, rho: float = 7.0) -> np.ndarray:
    ramp = np.linspace(0, 1, num_steps)
    min_inv, max_inv = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (max_inv + ramp * (min_inv - max_inv)) ** rho
    return np.concatenate([sig, [0.0]]).astype(np.float32)


def arch_invariant_rand(shape, seed: int) -> np.ndarray:
    """utils/misc.py:133-154: numpy RandomState(seed).standard_normal(shape) as float32."""
    return np.random.RandomState(seed).standard_normal(shape).astype(np.float32)


def bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def denoise_step(net, xt, gt, noise, indicator_t, sigma: float, sigma_next: float, guidance: float,
                 sigma_data: float = 0.5, sigma_aug: float = 0.001, return_net_output: bool = False):
    """model_v2w.py:130-149 for one step.  net(x_in, timestep, cond: bool) -> [16,T,H,W] (fp32).
    xt, gt [16,T,H,W]; noise fp32; indicator_t [T].  bf16 storage of x~, x_in, net outputs, x_next as in
    the reference's bf16 tensors; everything else fp32."""
    ind = indicator_t.float().reshape(1, -1, 1, 1)
    if sigma_aug >= sigma:
        ind = torch.zeros_like(ind)
    aug = (gt.float() + noise.float() * sigma_aug) / math.sqrt(sigma_aug ** 2 + sigma_data ** 2)
    aug = aug * math.sqrt(sigma ** 2 + sigma_data ** 2)
    xs = bf16(ind * aug + (1 - ind) * xt.float())
    x_in = bf16(xs / math.sqrt(sigma ** 2 + sigma_data ** 2))
    t = float(bf16(torch.tensor(0.25 * math.log(sigma), dtype=torch.float32)))
    oc = bf16(net(x_in, t, True))
    ou = bf16(net(x_in, t, False))
    o = oc + guidance * (oc - ou)
    net_output = o
    c_skip = sigma_data ** 2 / (sigma ** 2 + sigma_data ** 2)
    c_out = sigma * sigma_data / math.sqrt(sigma ** 2 + sigma_data ** 2)
    lat = (gt.float() - c_skip * xs) / c_out
    o = ind * lat + (1 - ind) * o
    x0 = c_skip * xs + c_out * o
    d = (xs - x0) / sigma
    nxt = bf16(xs + d * (sigma_next - sigma))
    return (nxt, net_output) if return_net_output else nxt
