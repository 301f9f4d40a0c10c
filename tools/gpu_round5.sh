#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe5.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe5.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe5.log; }
TO=300 run ops5 python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k attention
TO=300 run dit5 python -m pytest tests/test_dit_gpu.py -q --no-header -p no:cacheprovider -s
for p in 0 8 4 2; do
  G3C_ATTN_POLY=$p TO=300 TAILN=4 run perf_attn5_poly$p python tools/gpu_perf.py attn
done
TO=900 TAILN=3 run bench5 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
