#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe4.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe4.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe4.log; }
TO=300 run ops4 python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider
TO=300 run dit4 python -m pytest tests/test_dit_gpu.py -q --no-header -p no:cacheprovider -s
TO=400 run perf_gemm4 python tools/gpu_perf.py gemm
for p in 0 8 4 2; do
  G3C_ATTN_POLY=$p TO=300 TAILN=4 run perf_attn_poly$p python tools/gpu_perf.py attn
done
