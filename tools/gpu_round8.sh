#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe8.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe8.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe8.log; }
TO=300 run ops8 python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider
for p in 0 8 4; do
  G3C_ATTN_POLY=$p TO=300 TAILN=4 run perf8_poly$p python tools/gpu_perf.py attn
done
G3C_ATTN_IMPL=v1 G3C_ATTN_POLY=0 TO=300 TAILN=4 run perf8_v1 python tools/gpu_perf.py attn
