"""Minimal launch targets for `ncu --set full`: one self-attention, one projection GEMM, one gated GEMM and one
cache render at BASELINE sizes.  Usage: python tools/ncu_target.py [attn|gemm|warp]"""
import sys

import torch

sys.path.insert(0, ".")
from gen3c_b200 import ops, warp  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "attn"


def bf(*s, sc=1.0):
    return (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16)


if which == "attn":
    L, H = 56320, 32
    # the engine's calling convention: softmax scale * log2(e) folded into Q, scale = ln 2
    q, k, vt = bf(L, H * 128, sc=128 ** -0.5 * 1.4426950408889634), bf(L, H * 128), bf(H * 128, L)
    for _ in range(2):
        o = ops.attention(q, k, vt, H, scale=0.6931471805599453)
elif which == "sdpa":
    # torch's fused SDPA (cuDNN / flash backend) on the same shape: what the hot kernel is measured against
    L, H = 56320, 32
    q, k, v = bf(1, H, L, 128, sc=0.3), bf(1, H, L, 128), bf(1, H, L, 128)
    for _ in range(2):
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
elif which == "attn_cp8":
    L, H = 56320, 32
    q, k, vt = bf(L // 8, H * 128, sc=128 ** -0.5 * 1.4426950408889634), bf(L, H * 128), bf(H * 128, L)
    for _ in range(2):
        o = ops.attention(q, k, vt, H, scale=0.6931471805599453)
elif which == "eltwise":
    L, D = 56320, 4096
    x = torch.randn(L, D, device="cuda")
    pos = bf(L, D)
    sh, sc_ = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    for _ in range(2):
        ops.ln_modulate(x, sh, sc_, pos=pos)
    a, w = bf(512, 1024), bf(4096, 1024, sc=0.02)           # context K projection (1-CTA kernel)
    gq = torch.ones(128, device="cuda")
    for _ in range(2):
        ops.gemm_norm_rope(a, w, gq)
        ops.gemm(bf(L, 4096), bf(64, 4096, sc=0.02), ops.EPI_F32, block_n=64)   # final layer
elif which == "gemm":
    L = 56320
    a, w = bf(L, 4096), bf(4096, 4096, sc=0.02)
    x = torch.zeros(L, 4096, device="cuda")
    gate = torch.ones(4096, device="cuda")
    for _ in range(2):
        ops.gemm(a, w, ops.EPI_BF16)
        ops.gemm(a, w, ops.EPI_GATED_RESIDUAL_F32, out=x, gate=gate)
elif which == "warp":
    from oracle import cases

    h, w_, F = 704, 1280, 8
    depth = torch.from_numpy(cases.smooth_depth(h, w_)[None, None]).cuda()
    K = torch.from_numpy(cases.intrinsics(h, w_)[None]).cuda()
    eye = torch.eye(4, device="cuda")[None]
    img = torch.rand(1, 3, h, w_, device="cuda") * 2 - 1
    pts = warp.unproject_points(depth, eye, K)
    w2cs = torch.from_numpy(cases.pan_trajectory(F, 0.3)).cuda()[None]
    Ks = K[None].expand(1, F, 3, 3).contiguous()
    for _ in range(2):
        warp.render_cache(pts[None, None], img[None, None], None, w2cs, Ks, max_items_per_pass=4)
torch.cuda.synchronize()
