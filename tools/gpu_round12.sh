#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe12.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe12.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe12.log; }
TO=600 TAILN=8 run alltests12 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider
TO=900 TAILN=3 run bench12 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
G3C_GEMM_2CTA=0 TO=900 TAILN=3 run bench12_1cta python bench.py --steps 2 --warmup 3 --no-cpu-baseline
