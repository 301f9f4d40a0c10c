#!/bin/bash
mkdir -p gpurun_out
export G3C_PERF_FAST=1 G3C_PERF_LOG2=1
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe18.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe18.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe18.log; }
TO=200 TAILN=12 run tests18 python -m pytest tests/test_dit_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k attention --durations=5
G3C_ATTN_CLUSTER=1 TO=150 TAILN=12 run tests18_cl python -m pytest tests/test_dit_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k attention -x
TO=120 TAILN=4 run perf18_m2p python tools/gpu_perf.py attn
G3C_ATTN_CLUSTER=1 TO=120 TAILN=4 run perf18_clp python tools/gpu_perf.py attn
TO=120 TAILN=4 run perf18_m2pb python tools/gpu_perf.py attn
G3C_ATTN_CLUSTER=1 TO=120 TAILN=4 run perf18_clpb python tools/gpu_perf.py attn
G3C_ATTN_CLUSTER=1 TO=300 TAILN=3 run bench18_cl python bench.py --steps 2 --warmup 3 --no-cpu-baseline
