#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe6.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe6.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe6.log; }
TO=300 run ops6 python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k attention
G3C_ATTN_V2=1 TO=300 run ops6v2 python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k attention
for p in 0 8 4; do
  G3C_ATTN_POLY=$p TO=300 TAILN=4 run perf6_v1_poly$p python tools/gpu_perf.py attn
done
for p in 0 4; do
  G3C_ATTN_V2=1 G3C_ATTN_POLY=$p TO=300 TAILN=4 run perf6_v2_poly$p python tools/gpu_perf.py attn
done
G3C_ATTN_V2=1 G3C_ATTN_POLY=0 TO=600 TAILN=3 run ncu_attn_v2 ncu --set full --clock-control none --import-source on -k regex:k_attn_fwd -s 1 -c 1 -o gpurun_out/r01_attn_v2 -f python tools/ncu_target.py attn
G3C_ATTN_POLY=4 TO=600 TAILN=3 run ncu_attn_v1b ncu --set full --clock-control none --import-source on -k regex:k_attn_fwd -s 1 -c 1 -o gpurun_out/r01_attn_v1b -f python tools/ncu_target.py attn
