#!/bin/bash
# First-contact GPU battery: each group in its own process under a timeout so that a trap or a hang in
# one kernel cannot take the others down.  Output lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe.log; }
TO=300 run warp python -m pytest tests/test_warp_gpu.py -q -x --no-header -p no:cacheprovider
TO=300 run ops_gemm python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k "gemm"
TO=300 run ops_attn python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k "attention"
TO=200 run ops_elt python -m pytest tests/test_dit_ops_gpu.py -q --no-header -p no:cacheprovider -k "ln_modulate or rmsnorm"
TO=300 run dit python -m pytest tests/test_dit_gpu.py -q --no-header -p no:cacheprovider -s
TO=300 run perf_warp python tools/gpu_perf.py warp eltwise
TO=400 run perf_gemm python tools/gpu_perf.py gemm
TO=400 run perf_attn python tools/gpu_perf.py attn
