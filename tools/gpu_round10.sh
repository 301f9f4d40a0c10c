#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe10.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe10.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe10.log; }
G3C_ATTN_IMPL=v5 TO=300 run ops10 python -m pytest tests/test_dit_ops_gpu.py tests/test_fullsize_properties_gpu.py -q --no-header -p no:cacheprovider -k attention
for p in 0 8 4 2; do
  G3C_ATTN_IMPL=v5 G3C_ATTN_POLY=$p TO=300 TAILN=4 run perf10_v5_poly$p python tools/gpu_perf.py attn
done
G3C_ATTN_IMPL=v5 TO=300 TAILN=8 run trace_v5 python tools/attn_trace.py
