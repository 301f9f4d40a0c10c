#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/quick.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/quick.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/quick.log; }
TO=200 TAILN=8 run fuse_unit python -m pytest tests/test_dit_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "gemm"
G3C_FUSE_NORM_ROPE=1 TO=200 TAILN=8 run fuse_engine python -m pytest tests/test_dit_gpu.py tests/test_fullsize_properties_gpu.py -q -m gpu --no-header -p no:cacheprovider
G3C_FUSE_NORM_ROPE=1 TO=200 TAILN=3 run bench_fuse python bench.py --steps 2 --warmup 3 --no-cpu-baseline
G3C_FUSE_NORM_ROPE=0 TO=200 TAILN=3 run bench_nofuse python bench.py --steps 2 --warmup 3 --no-cpu-baseline
