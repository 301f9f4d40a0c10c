"""Thin torch-tensor wrappers over the per-operator C ABI entry points of Path D (used by the tests,
and usable as drop-in operators, e.g. ``attention`` as an ``attn_op`` for the reference's
``Attention(attn_op=...)`` seam, module/attention.py:136-139).  CUDA only; no fallback."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

EPI_BF16, EPI_GELU_BF16, EPI_GATED_RESIDUAL_F32, EPI_F32 = 0, 1, 2, 3


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {dtype}")


def gemm(a: torch.Tensor, b: torch.Tensor, epilogue: int = EPI_BF16, out: Optional[torch.Tensor] = None,
         gate: Optional[torch.Tensor] = None, block_n: int = 0) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[N,K]^T on tcgen05 (bf16 in, fp32 accumulate).
    epilogue: EPI_BF16 | EPI_GELU_BF16 (bf16 out) | EPI_F32 (f32 out) | EPI_GATED_RESIDUAL_F32 (out f32 += gate*acc)."""
    _chk(a, torch.bfloat16, "a")
    _chk(b, torch.bfloat16, "b")
    M, K = a.shape
    N, K2 = b.shape
    assert K == K2, "inner dimensions differ"
    odt = torch.bfloat16 if epilogue in (EPI_BF16, EPI_GELU_BF16) else torch.float32
    if out is None:
        if epilogue == EPI_GATED_RESIDUAL_F32:
            raise ValueError("gated-residual epilogue accumulates into `out`")
        out = torch.empty((M, N), device=a.device, dtype=odt)
    _chk(out, odt, "out")
    if gate is not None:
        _chk(gate, torch.float32, "gate")
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.g3c_gemm_bf16(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), M, N, K, K, K, N, epilogue,
                                     _lib.ptr(gate), block_n, _lib.stream_ptr()), "g3c_gemm_bf16")
    return out


def gemm_norm_rope(a: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, cos_sin: Optional[torch.Tensor] = None,
                   eps: float = 1e-6) -> torch.Tensor:
    """out[M,N] (bf16) = RoPE(RMSNorm_head(a @ b^T) * gamma): projection + per-head (128) RMSNorm + rotate-half RoPE in
    one kernel (reference: to_q / to_k = Sequential(Linear, RMSNorm) + apply_rotary_pos_emb, module/attention.py:263-283).
    gamma f32 [128]; cos_sin f32 [M, 128] (cos | sin of the 64 angles) or None."""
    _chk(a, torch.bfloat16, "a")
    _chk(b, torch.bfloat16, "b")
    _chk(gamma, torch.float32, "gamma")
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and N % 128 == 0 and gamma.numel() == 128
    if cos_sin is not None:
        _chk(cos_sin, torch.float32, "cos_sin")
        assert tuple(cos_sin.shape) == (M, 128)
    out = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.g3c_gemm_norm_rope_bf16(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), M, N, K, K, K, N, _lib.ptr(gamma),
                                               _lib.ptr(cos_sin), eps, _lib.stream_ptr()), "g3c_gemm_norm_rope_bf16")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, heads: int, scale: Optional[float] = None,
              vt_chunk_len: int = 0) -> torch.Tensor:
    """q [Lq, heads*128], k [Lk, heads*128], vt [chunks, heads*128, chunk_len] or [heads*128, Lk] (V transposed).
    Returns o [Lq, heads*128] = softmax(q k^T * scale) v per head."""
    for t, n in ((q, "q"), (k, "k"), (vt, "vt")):
        _chk(t, torch.bfloat16, n)
    Lq, Dq = q.shape
    Lk = k.shape[0]
    assert Dq == heads * 128 and k.shape[1] == heads * 128
    o = torch.empty_like(q)
    if scale is None:
        scale = 128 ** -0.5
    lib = _lib.load()
    with torch.cuda.device(q.device):
        _lib.check(lib.g3c_attn_fwd(_lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(o), Lq, Lk, heads, Dq, Dq, Dq,
                                    vt_chunk_len, scale, _lib.stream_ptr()), "g3c_attn_fwd")
    return o


def ln_modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, pos: Optional[torch.Tensor] = None,
                eps: float = 1e-6) -> torch.Tensor:
    """x (f32 [L,D], updated in place when pos is given: x += pos) ; returns bf16 LN(x)*(1+scale)+shift."""
    _chk(x, torch.float32, "x")
    L, D = x.shape
    y = torch.empty((L, D), device=x.device, dtype=torch.bfloat16)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.g3c_ln_modulate(_lib.ptr(x), _lib.ptr(pos), _lib.ptr(shift), _lib.ptr(scale), _lib.ptr(y), L, D,
                                       eps, _lib.stream_ptr()), "g3c_ln_modulate")
    return y


def rmsnorm_rope_(qk: torch.Tensor, heads: int, gamma: torch.Tensor, cos_sin: Optional[torch.Tensor] = None,
                  eps: float = 1e-6) -> torch.Tensor:
    """In place: per-head RMSNorm (+ rotate-half RoPE with cos_sin [L,128] = cos|sin of the 64 angles)."""
    _chk(qk, torch.bfloat16, "qk")
    _chk(gamma, torch.float32, "gamma")
    L, ld = qk.shape
    lib = _lib.load()
    with torch.cuda.device(qk.device):
        _lib.check(lib.g3c_rmsnorm_rope(_lib.ptr(qk), ld, L, heads, _lib.ptr(gamma), _lib.ptr(cos_sin), eps,
                                        _lib.stream_ptr()), "g3c_rmsnorm_rope")
    return qk
