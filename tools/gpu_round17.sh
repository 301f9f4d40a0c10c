#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe17.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe17.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe17.log; }
TO=400 TAILN=15 run tests17 python -m pytest tests/test_dit_ops_gpu.py tests/test_dit_gpu.py -q -m gpu --no-header -p no:cacheprovider
G3C_ATTN_MODE=0 TO=300 TAILN=6 run perf17_m0 python tools/gpu_perf.py attn
TO=300 TAILN=6 run perf17_m2 python tools/gpu_perf.py attn
G3C_PERF_LOG2=1 TO=300 TAILN=6 run perf17_m2p python tools/gpu_perf.py attn
G3C_ATTN_MODE=0 TO=300 TAILN=6 run perf17_m0b python tools/gpu_perf.py attn
TO=300 TAILN=6 run perf17_m2b python tools/gpu_perf.py attn
G3C_PERF_LOG2=1 TO=300 TAILN=6 run perf17_m2pb python tools/gpu_perf.py attn
TO=900 TAILN=3 run bench17 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
