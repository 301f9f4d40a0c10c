#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe7.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe7.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe7.log; }
TO=300 run ops7 python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k attention
G3C_ATTN_IMPL=v1 TO=300 run ops7v1 python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k attention
for p in 0 8 4; do
  G3C_ATTN_POLY=$p TO=300 TAILN=4 run perf7_poly$p python tools/gpu_perf.py attn
done
G3C_ATTN_IMPL=v1 G3C_ATTN_POLY=0 TO=300 TAILN=4 run perf7_v1 python tools/gpu_perf.py attn
G3C_ATTN_POLY=0 TO=600 TAILN=3 run ncu_attn_v4 ncu --set full --clock-control none --import-source on -k regex:k_attn_fwd -s 1 -c 1 -o gpurun_out/r01_attn_v4 -f python tools/ncu_target.py attn
