#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe3.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe3.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe3.log; }
TO=300 TAILN=30 run warp3 python -m pytest tests/test_warp_gpu.py -q --no-header -p no:cacheprovider
TO=900 TAILN=5 run bench3 python bench.py --steps 2 --warmup 3
TO=600 TAILN=5 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:k_attn_fwd -s 1 -c 1 -o gpurun_out/r01_attn -f python tools/ncu_target.py attn
TO=600 TAILN=5 run ncu_gemm ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 2 -c 2 -o gpurun_out/r01_gemm -f python tools/ncu_target.py gemm
TO=600 TAILN=5 run ncu_warp ncu --set full --clock-control none --import-source on -k regex:k_splat_points -s 2 -c 1 -o gpurun_out/r01_splat -f python tools/ncu_target.py warp
TO=900 TAILN=5 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1500 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline
