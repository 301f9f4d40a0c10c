#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe11.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe11.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe11.log; }
TO=200 TAILN=30 run gemm2_test python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k "cta_pair" -x
TO=300 TAILN=12 run gemm2_perf python tools/gpu_perf.py gemm
