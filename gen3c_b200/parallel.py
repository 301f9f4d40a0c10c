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


def _robust_broadcast(tensor: torch.Tensor, src: int, pg) -> torch.Tensor:
    """reference parallel.py:136-163: broadcast the shape first so that receivers may hold a tensor of any shape."""
    dev = tensor.device
    if dist.get_rank() == src:
        shape = torch.tensor(tensor.shape, device=dev)
    else:
        shape = torch.empty(tensor.dim(), dtype=torch.long, device=dev)
    dist.broadcast(shape, src, group=pg)
    if dist.get_rank() != src:
        tensor = tensor.new_empty(shape.tolist()).type_as(tensor)
    tensor = tensor.contiguous()
    dist.broadcast(tensor, src, group=pg)
    return tensor


def broadcast(item, cp_group=None):
    """reference parallel.py:90-133 for the context-parallel group (there is no tensor parallelism in GEN3C
    inference): the item of the group's lowest global rank replaces everybody's."""
    if cp_group is None or not dist.is_initialized() or dist.get_world_size(cp_group) <= 1:
        return item
    src = min(dist.get_process_group_ranks(cp_group))
    if isinstance(item, torch.Tensor):
        return _robust_broadcast(item, src, cp_group)
    if item is not None:
        box = [item]
        dist.broadcast_object_list(box, src, group=cp_group)
        item = box[0]
    return item


def broadcast_condition(condition, to_tp: bool = True, to_cp: bool = True, cp_group=None):
    """reference model_v2w.py broadcast_condition: every tensor / picklable field of the condition object is replaced by
    rank-min's copy, so that all context-parallel ranks denoise against identical conditions."""
    if not to_cp or cp_group is None:
        return condition
    for key, value in list(vars(condition).items()):
        if key == "extra":
            continue
        setattr(condition, key, broadcast(value, cp_group))
    return condition
