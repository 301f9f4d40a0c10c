"""Context-parallel host logic on CPU: world_size 2, gloo.  Each rank owns a contiguous slice of the latent
frames (module/parallel.py:44-53), runs the oracle network on its slice and exchanges K/V with an all-gather —
the same partition, offsets and gather order the CUDA engine uses (k_all = [rank][L_local][D]).  The gathered
result must equal the single-process forward (reference semantics: general_dit.py:524-543)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cases, dit_oracle


def _worker(rank, world, port, T, H, W, M, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gen3c_b200.parallel import split_inputs_cp, cat_outputs_cp

        cfg = cases.TINY
        sd = dit_oracle.random_state_dict(cfg, seed=3)
        inp = cases.dit_inputs(cfg, T, H, W, M, seed=5)
        Tl = T // world

        def gather(i, k, v):
            ks = [torch.empty_like(k) for _ in range(world)]
            vs = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(ks, k.contiguous())
            dist.all_gather(vs, v.contiguous())
            return torch.cat(ks), torch.cat(vs)

        xs = split_inputs_cp(inp["x"][None], seq_dim=2, cp_group=dist.group.WORLD)[0]
        ms = split_inputs_cp(inp["cond_mask"][None], seq_dim=2, cp_group=dist.group.WORLD)[0]
        ps = split_inputs_cp(inp["pose"][None], seq_dim=2, cp_group=dist.group.WORLD)[0]
        out = dit_oracle.forward(sd, cfg, xs, ms, ps, inp["padding"], inp["timestep"], inp["ctx_c"], t0=rank * Tl,
                                 kv_gather=gather)
        full = cat_outputs_cp(out[None].contiguous(), seq_dim=2, cp_group=dist.group.WORLD)[0]
        if rank == 0:
            ret.put(full)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_context_parallel_gloo_world2_equals_single():
    T, H, W, M = 4, 16, 16, 128
    cfg = cases.TINY
    sd = dit_oracle.random_state_dict(cfg, seed=3)
    inp = cases.dit_inputs(cfg, T, H, W, M, seed=5)
    want = dit_oracle.forward(sd, cfg, inp["x"], inp["cond_mask"], inp["pose"], inp["padding"], inp["timestep"], inp["ctx_c"])
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, H, W, M, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert float((got - want).norm() / want.norm()) < 1e-5


def test_split_requires_divisibility():
    from gen3c_b200.parallel import chunk_bounds

    assert chunk_bounds(16, 3, 8) == (6, 2)
    with pytest.raises(AssertionError):
        chunk_bounds(16, 0, 3)
