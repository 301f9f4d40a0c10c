#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe_cp4c.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe_cp4c.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe_cp4c.log; }
TO=200 TAILN=3 run bench_n4_p2p2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 2 --warmup 3 --cp-mode p2p --no-cpu-baseline
