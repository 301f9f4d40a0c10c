#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe13.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe13.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe13.log; }
TO=300 TAILN=6 run attntests13 python -m pytest tests/test_dit_ops_gpu.py tests/test_dit_gpu.py -q -m gpu --no-header -p no:cacheprovider
TO=300 TAILN=6 run perf13_half python tools/gpu_perf.py attn
G3C_ATTN_PHALF=0 TO=300 TAILN=6 run perf13_full python tools/gpu_perf.py attn
TO=300 TAILN=40 run trace13 python tools/attn_trace.py
TO=900 TAILN=3 run bench13 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
