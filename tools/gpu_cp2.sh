#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe_cp.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe_cp.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe_cp.log; }
nvidia-smi -L | tee -a gpurun_out/probe_cp.log
TO=400 TAILN=30 run cp_test python -m pytest tests/test_cp_gpu.py -q --no-header -p no:cacheprovider -s
TO=900 TAILN=6 run bench_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3
