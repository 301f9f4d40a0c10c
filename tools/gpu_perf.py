"""Kernel-level timings at BASELINE sizes (CUDA events, warm-up, L2 flushed between iterations).
Usage: python tools/gpu_perf.py [gemm] [attn] [warp] [eltwise]  -> JSON lines on stdout."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from gen3c_b200 import ops, warp  # noqa: E402

flush = None


def timeit(fn, iters=5, warm=2):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bf(*shape, s=1.0):
    return (torch.randn(*shape, device="cuda") * s).to(torch.bfloat16)


def gemm_cases():
    L = 56320
    for name, M, N, K, epi in [("qkv/out 4096x4096", L, 4096, 4096, ops.EPI_BF16),
                               ("vT swapped", 4096, L, 4096, ops.EPI_BF16),
                               ("mlp1 gelu", L, 16384, 4096, ops.EPI_GELU_BF16),
                               ("mlp2 gated", L, 4096, 16384, ops.EPI_GATED_RESIDUAL_F32),
                               ("out gated", L, 4096, 4096, ops.EPI_GATED_RESIDUAL_F32),
                               ("L/8 4096x4096", L // 8, 4096, 4096, ops.EPI_BF16)]:
        a, b = bf(M, K), bf(N, K, s=0.02)
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi >= 2 else torch.bfloat16)
        gate = torch.randn(N, device="cuda")
        med, best = timeit(lambda: ops.gemm(a, b, epi, out=out, gate=gate if epi == 2 else None))
        fl = 2.0 * M * N * K
        pair_ms = None
        if N % 256 == 0:
            pair_ms, _ = timeit(lambda: ops.gemm(a, b, epi, out=out, gate=gate if epi == 2 else None, block_n=512))
        tmed, _ = timeit(lambda: torch.matmul(a, b.T))
        print(json.dumps({"kernel": "gemm", "case": name, "M": M, "N": N, "K": K, "ms": med, "ms_best": best,
                          "tflops": fl / med / 1e9, "pair_ms": pair_ms, "pair_tflops": (fl / pair_ms / 1e9) if pair_ms else None, "cublas_ms": tmed, "cublas_tflops": fl / tmed / 1e9}), flush=True)
        del a, b, out


def attn_cases():
    for name, Lq, Lk, heads in [("self 56320 x 56320 (32 heads)", 56320, 56320, 32),
                                ("self cp8 7040 x 56320", 7040, 56320, 32),
                                ("cross 56320 x 512", 56320, 512, 32)]:
        D = heads * 128
        q, k, vt = bf(Lq, D), bf(Lk, D), bf(D, Lk)
        # G3C_PERF_LOG2=1: the engine's calling convention (softmax scale * log2 e folded into Q, scale = ln 2)
        log2u = os.environ.get("G3C_PERF_LOG2", "0") == "1"
        if log2u:
            q = (q.float() * (128 ** -0.5 * 1.4426950408889634)).to(torch.bfloat16)
        sc = 0.6931471805599453 if log2u else None
        med, best = timeit(lambda: ops.attention(q, k, vt, heads, scale=sc), iters=3, warm=1)
        fl = 4.0 * Lq * Lk * D
        rec = {"kernel": "attn", "case": name, "ms": med, "ms_best": best, "tflops": fl / med / 1e9}
        try:
            if os.environ.get("G3C_PERF_FAST", "0") == "1":
                raise RuntimeError("skipped")
            qh = q.reshape(Lq, heads, 128).permute(1, 0, 2)[None]
            kh = k.reshape(Lk, heads, 128).permute(1, 0, 2)[None]
            vh = vt.T.reshape(Lk, heads, 128).permute(1, 0, 2)[None].contiguous()
            tm, _ = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh), iters=3, warm=1)
            rec.update({"torch_sdpa_ms": tm, "torch_sdpa_tflops": fl / tm / 1e9})
        except Exception as ex:  # noqa: BLE001
            rec["torch_sdpa_error"] = str(ex)[:100]
        print(json.dumps(rec), flush=True)
        del q, k, vt


def warp_cases():
    from oracle import cases
    import numpy as np

    h, w, F = 704, 1280, 121
    depth = torch.from_numpy(cases.smooth_depth(h, w)[None, None]).cuda()
    K = torch.from_numpy(cases.intrinsics(h, w)[None]).cuda()
    eye = torch.eye(4, device="cuda")[None]
    img = torch.rand(1, 3, h, w, device="cuda") * 2 - 1
    pts = warp.unproject_points(depth, eye, K)
    w2cs = torch.from_numpy(cases.pan_trajectory(F, 0.3)).cuda()[None]
    Ks = K[None].expand(1, F, 3, 3).contiguous()
    for items in (2, 4, 8):
        med, best = timeit(lambda: warp.render_cache(pts[None, None], img[None, None], None, w2cs, Ks,
                                                     max_items_per_pass=items), iters=5, warm=2)
        algo = 44.0 * h * w * F
        print(json.dumps({"kernel": "render_cache", "case": f"121 frames 704x1280 N=1 items/pass={items}", "ms": med,
                          "ms_best": best, "frames_per_s": F / med * 1e3, "algo_GBps": algo / med / 1e6}), flush=True)


def eltwise_cases():
    L, D = 56320, 4096
    x = torch.randn(L, D, device="cuda")
    pos = bf(L, D)
    sh, sc = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    med, _ = timeit(lambda: ops.ln_modulate(x, sh, sc, pos=pos))
    print(json.dumps({"kernel": "ln_modulate+pos", "ms": med, "GBps": (L * D * (4 + 4 + 2 + 2)) / med / 1e6}), flush=True)
    q = bf(L, D)
    gamma = torch.ones(128, device="cuda")
    cs = torch.rand(L, 128, device="cuda")
    med, _ = timeit(lambda: ops.rmsnorm_rope_(q, 32, gamma, cs))
    print(json.dumps({"kernel": "rmsnorm_rope", "ms": med, "GBps": (L * D * 4) / med / 1e6}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attn", "warp", "eltwise"]
    for wname in which:
        {"gemm": gemm_cases, "attn": attn_cases, "warp": warp_cases, "eltwise": eltwise_cases}[wname]()
