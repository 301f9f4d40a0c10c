"""Context-parallel helpers with the reference's names and semantics (Path D row D11).

reference: cosmos_predict1/diffusion/module/parallel.py — split_inputs_cp :25-53, cat_outputs_cp :56-87.
The split is contiguous along `seq_dim` into `cp_size` equal chunks; rank r owns chunk r."""
from __future__ import annotations

import torch
import torch.distributed as dist


def chunk_bounds(n: int, rank: int, size: int) -> tuple[int, int]:
    """(start, length) of rank's chunk; the reference asserts divisibility (parallel.py:47)."""
    assert n % size == 0, f"sequence length {n} is not divisible by the context-parallel size {size}"
    c = n // size
    return rank * c, c


def split_inputs_cp(x: torch.Tensor, seq_dim: int, cp_group) -> torch.Tensor:
    rank, size = dist.get_rank(cp_group), dist.get_world_size(cp_group)
    start, length = chunk_bounds(x.shape[seq_dim], rank, size)
    return x.narrow(seq_dim, start, length).contiguous()


def cat_outputs_cp(x: torch.Tensor, seq_dim: int, cp_group) -> torch.Tensor:
    size = dist.get_world_size(cp_group)
    parts = [torch.empty_like(x) for _ in range(size)]
    dist.all_gather(parts, x.contiguous(), group=cp_group)
    return torch.cat(parts, dim=seq_dim)
