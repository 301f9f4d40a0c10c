#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe2.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe2.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe2.log; }
TO=120 run dbg_warp python tools/debug_warp.py
TO=300 TAILN=40 run warp2 python -m pytest tests/test_warp_gpu.py -q --no-header -p no:cacheprovider
TO=300 run ops_gemm2 python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k "gemm"
TO=300 run dit2 python -m pytest tests/test_dit_gpu.py -q --no-header -p no:cacheprovider -s
TO=300 run smoke python __graft_entry__.py smoke
TO=400 run perf_gemm2 python tools/gpu_perf.py gemm
TO=900 TAILN=5 run bench python bench.py --steps 2 --warmup 3
TO=300 TAILN=5 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
