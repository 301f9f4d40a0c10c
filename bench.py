#!/usr/bin/env python
"""bench.py — denoise-steps/sec of the Cosmos-7B GEN3C DiT (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

One step = one loop body of generate_samples_from_batch (reference model_v2w.py:130-149): sampler glue +
cond forward + uncond forward of the 28-block 7B DiT over the 121-frame / 704x1280 latent [16,16,88,160]
(56 320 tokens), bf16 weights/activations, fp32 accumulation, random-init weights in the real checkpoint
layout, synthetic latents / poses / text context (no network for checkpoints or data).
N > 1 (torchrun): context parallel over the 16 latent frames (16/N per rank), one K and one V^T
all-gather per self-attention layer; total work is fixed -> "scaling": "strong".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FORWARD = 2.2096e15  # SURVEY.md §8d / BASELINE.md §2
FLOP_PER_STEP = 2 * FLOP_PER_FORWARD
SELF_ATTN_FLOP_PER_LAUNCH_FULL = 4.0 * 56320 * 56320 * 4096  # 5.197e13 at cp = 1
LAT = (16, 16, 88, 160)
CTX = (512, 1024)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"], "hbm_gbs": d["hbm_gbs"],
                "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_sample_seconds(threads: int, repeats: int = 1):
    """One FA-CA-MLP block at full width (D=4096, 32 heads, ctx 512x1024) on ONE latent frame (3 520 tokens),
    fp32, torch CPU ops = the reference's graph with the TE ops restated (oracle/dit_oracle.py).
    Returns (seconds per sample, FLOPs of the sample)."""
    import torch

    from oracle import dit_oracle

    torch.set_num_threads(threads)
    cfg = dit_oracle.DitCfg(num_blocks=1)
    g = torch.Generator().manual_seed(0)
    sd = {k: (0.02 * torch.randn(s, generator=g)) for k, s in dit_oracle.state_dict_shapes(cfg).items()}
    T, H, W = 1, 88, 160
    x = torch.randn(16, T, H, W, generator=g)
    mask = torch.zeros(1, T, H, W)
    pose = torch.randn(64, T, H, W, generator=g)
    ctx = torch.randn(*CTX, generator=g)
    L, D = T * 44 * 80, 4096
    flops = 28.0 * L * D * D + 4.0 * L * L * D + 4.0 * 512 * 1024 * D + 4.0 * L * 512 * D + 2.0 * L * 328 * D + 2.0 * L * D * 64
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            dit_oracle.forward(sd, cfg, x, mask, pose, None, 0.5, ctx)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, flops


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 64)  # torch's CPU GEMM/SDPA stop scaling (and regress) beyond this
    cpu_sample_seconds(threads)  # page-in / warm
    for _ in range(max(0, args.warmup - 1)):
        cpu_sample_seconds(threads)
    ts = []
    flops = 0.0
    for _ in range(args.steps):
        dt, flops = cpu_sample_seconds(threads)
        ts.append(dt)
    t = sum(ts) / len(ts)
    sps = 1.0 / (t * FLOP_PER_STEP / flops)
    sample = ("1 FA-CA-MLP block of the 7B DiT (D=4096, 32 heads, ctx 512x1024) on 1 latent frame (3 520 tokens), fp32 "
              "torch-CPU oracle port; steps/s extrapolated by FLOPs (x%.0f) to the full 2-forward step" % (FLOP_PER_STEP / flops))
    line = {"impl": "reference", "metric": "denoise-steps/sec, 7B DiT, 121-frame 720p latent", "value": sps,
            "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / sps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Cosmos-7B GEN3C DiT denoise step (2 forwards), latent [16,16,88,160], ctx 512x1024",
                       "note": "CPU-extrapolated; the reference has no CPU path of its own (model_t2w.py:56)"},
            "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def build_net(torch, device):
    from gen3c_b200.dit import VideoExtendGeneralDIT

    net = VideoExtendGeneralDIT(device=device)  # GEN3C_Cosmos_7B defaults
    g = torch.Generator(device=device).manual_seed(1234)
    with torch.no_grad():
        for k, p in net.state_dict(keep_vars=True).items():
            if k == "pos_embedder.seq":
                continue
            if p.dim() == 1:
                p.copy_((1.0 + 0.05 * torch.randn(p.shape, device=device, generator=g)).to(p.dtype))
            else:
                p.copy_((0.02 * torch.randn(p.shape, device=device, generator=g)).to(p.dtype))
    return net


def run_ours(args):
    import torch
    import torch.distributed as dist

    from gen3c_b200 import _lib, sampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    net = build_net(torch, dev)
    if world > 1:
        assert LAT[1] % world == 0, "16 latent frames must divide over the ranks"
        net.enable_context_parallel(dist.group.WORLD, mode=args.cp_mode)
    Tl = LAT[1] // world
    t0 = rank * Tl
    bf = torch.bfloat16
    g = torch.Generator().manual_seed(1)  # host-side synthetic inputs (pinned), same on every rank
    sch = sampler.EDMEulerScheduler().set_timesteps(35)

    def pin(t):
        return t.contiguous().pin_memory()

    full = {
        "xt": (torch.randn(LAT, generator=g) * sch.init_noise_sigma).to(bf),
        "gt": (0.5 * torch.randn(LAT, generator=g)).to(bf),
        "noise": sampler.arch_invariant_rand(LAT, 1),
        "pose": (0.5 * torch.randn(64, *LAT[1:], generator=g)).to(bf),
        "mask": torch.zeros(1, *LAT[1:]).to(bf),
        "ctx_c": torch.randn(CTX, generator=g).to(bf),
        "ctx_u": torch.randn(CTX, generator=g).to(bf),
        "pad": torch.zeros(LAT[2], LAT[3]).to(bf),
    }
    full["mask"][:, 0] = 1
    ind_full = torch.zeros(LAT[1])
    ind_full[0] = 1.0
    host = {k: pin(v[:, t0:t0 + Tl] if v.dim() == 4 else v) for k, v in full.items()}
    host["ind"] = pin(ind_full[t0:t0 + Tl])
    devt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    out_host = torch.empty((16, Tl, LAT[2], LAT[3]), dtype=bf).pin_memory()
    sig = [float(s) for s in sch.sigmas]
    lib = _lib.load()

    def step_resident(i):
        return sampler.denoise_step(net, devt["xt"], devt["gt"], devt["noise"], devt["ind"], devt["mask"], devt["pose"],
                                    devt["pad"], devt["ctx_c"], devt["ctx_u"], sig[i % 34], sig[i % 34 + 1], 1.0)

    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in ("xt", "gt", "noise", "pose", "mask", "ctx_c", "ctx_u", "pad", "ind"))
    d2h_bytes = out_host.numel() * out_host.element_size()

    def step_e2e(i):
        d = {k: host[k].to(dev, non_blocking=True) for k in host}
        o = sampler.denoise_step(net, d["xt"], d["gt"], d["noise"], d["ind"], d["mask"], d["pose"], d["pad"], d["ctx_c"],
                                 d["ctx_u"], sig[i % 34], sig[i % 34 + 1], 1.0)
        out_host.copy_(o, non_blocking=True)
        return o

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for i in range(max(3, args.warmup)):
        step_resident(i)
    barrier()
    launches_per_step = net.last_launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    # -- timed region 1: inputs resident in HBM, profiling events on (per-category device time)
    _lib.check(lib.g3c_dit_profile(net._engine(), 1), "g3c_dit_profile")
    ms_resident = timed(step_resident, args.steps)
    cat_ms = (C.c_float * 6)()
    cat_n = (C.c_int * 6)()
    _lib.check(lib.g3c_dit_profile_read(net._engine(), cat_ms, cat_n, 6), "g3c_dit_profile_read")
    _lib.check(lib.g3c_dit_profile(net._engine(), 0), "g3c_dit_profile")
    # -- timed region 2: same step through the public API with pinned host buffers (H2D + D2H inside)
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.finish()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    names = ["gemm", "attn_self", "attn_cross", "eltwise", "comm", "vector"]
    breakdown = {n: {"ms_per_step": cat_ms[i] / args.steps, "launches_per_step": cat_n[i] // max(1, args.steps)}
                 for i, n in enumerate(names)}
    # dominant kernel: self-attention (65.9 % of the FLOPs); algorithmic FLOPs per launch on this rank
    attn_launches = max(1, cat_n[1])
    attn_ms = cat_ms[1] / attn_launches
    attn_flop = SELF_ATTN_FLOP_PER_LAUNCH_FULL / world
    achieved = attn_flop / (attn_ms * 1e-3) / 1e12 if attn_ms > 0 else 0.0
    peak = peaks["tflops_sustained"]
    sps = 1e3 / ms_resident
    line = {
        "metric": "denoise-steps/sec, 7B DiT, 121-frame 720p latent", "value": sps, "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_resident,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Cosmos-7B GEN3C DiT denoise step (cond + uncond forward + EDM Euler glue), latent "
                               "[16,16,88,160] = 56 320 tokens, ctx 512x1024, guidance 1, random-init weights",
                   "parallelism": f"cp{world}" + (f" ({args.cp_mode or os.environ.get('G3C_CP_MODE', 'p2p')} K/V exchange)" if world > 1 else ""), "l2": "inputs larger than L2 (14.5 GB weights, 0.9 GB residual stream)"},
        "e2e": {"value": 1e3 / ms_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "tensor", "kernel": "k_attn_fwd (self-attention)", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes(), "peak_source": peaks["source"] + " (sustained)",
                     "step_tflops": FLOP_PER_STEP * sps / 1e12 / world, "step_frac": FLOP_PER_STEP * sps / 1e12 / world / peak},
        "kernel_breakdown": breakdown,
        "clocks": clk,
        "workspace_gb": net.workspace_bytes() / 1e9,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = min(os.cpu_count() or 1, 64)
        cpu_sample_seconds(threads)
        dt, flops = cpu_sample_seconds(threads)
        line["cpu_baseline"] = {
            "value": 1.0 / (dt * FLOP_PER_STEP / flops), "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": "1 FA-CA-MLP block (D=4096) on 1 latent frame (3 520 tokens), fp32 torch-CPU oracle port, %.1f s; "
                      "extrapolated by FLOPs (x%.0f)" % (dt, FLOP_PER_STEP / flops)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic_bytes():
    """DRAM bytes (read + write) of one full-size self-attention launch, from the committed `ncu --set full` summary
    (profiles/r01_attn_ncu_summary.txt); None when the summary is absent.  Static evidence, not measured by this run."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_attn_ncu_summary.txt")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot, seen = 0.0, 0
    try:
        for ln in open(path):
            if ln.startswith("--") and seen >= 2:
                break
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if ln.startswith(key):
                    unit = ln[ln.index("[") + 1:ln.index("]")]
                    tot += float(ln.split("=")[1]) * scale[unit]
                    seen += 1
    except (OSError, ValueError, KeyError):
        return None
    return tot if seen >= 2 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cp-mode", default=None, choices=[None, "p2p", "nccl"],
                    help="context-parallel K/V exchange: p2p = fused projection -> peer-memory all-gather (default), nccl = ncclAllGather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
