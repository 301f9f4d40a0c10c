#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe15.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe15.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe15.log; }
TO=300 TAILN=15 run attntests15 python -m pytest tests/test_dit_ops_gpu.py tests/test_dit_gpu.py -q -m gpu --no-header -p no:cacheprovider
G3C_ATTN_MODE=0 TO=300 TAILN=6 run attntests15_m0 python -m pytest tests/test_dit_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k attention
G3C_ATTN_MODE=0 TO=300 TAILN=6 run perf15_m0 python tools/gpu_perf.py attn
TO=300 TAILN=6 run perf15_m2 python tools/gpu_perf.py attn
G3C_ATTN_MODE=0 TO=300 TAILN=6 run perf15_m0b python tools/gpu_perf.py attn
TO=300 TAILN=6 run perf15_m2b python tools/gpu_perf.py attn
TO=300 TAILN=12 run trace15 python tools/attn_trace.py
TO=900 TAILN=3 run bench15 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
