#!/bin/bash
mkdir -p gpurun_out
export G3C_PERF_LOG2=1
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/quick.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/quick.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/quick.log; }
TO=200 TAILN=6 run quick_tests python -m pytest tests/test_dit_ops_gpu.py tests/test_dit_gpu.py -q -m gpu --no-header -p no:cacheprovider
G3C_PERF_FAST=1 TO=120 TAILN=4 run quick_perf python tools/gpu_perf.py attn
