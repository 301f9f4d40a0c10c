#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe16.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe16.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe16.log; }
TO=300 TAILN=15 run attntests16 python -m pytest tests/test_dit_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k attention
G3C_ATTN_MODE=0 TO=300 TAILN=6 run perf16_m0 python tools/gpu_perf.py attn
TO=300 TAILN=6 run perf16_m2 python tools/gpu_perf.py attn
G3C_ATTN_MODE=0 TO=300 TAILN=6 run perf16_m0b python tools/gpu_perf.py attn
TO=300 TAILN=6 run perf16_m2b python tools/gpu_perf.py attn
G3C_ATTN_TRACE_MMA_ONLY=1 TO=300 TAILN=4 run trace16_m2 python tools/attn_trace.py
G3C_ATTN_MODE=0 TO=300 TAILN=8 run trace16_m0 python tools/attn_trace.py
