#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe_cpd.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe_cpd.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe_cpd.log; }
TO=120 TAILN=30 run issue_rates tools/microbench/issue_rates
TO=400 TAILN=12 run cp_test4 python -m pytest tests/test_cp_gpu.py -q --no-header -p no:cacheprovider -s
TO=500 TAILN=3 run bench_n2_ce python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --cp-mode p2p
TO=500 TAILN=3 run bench_n2_nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --cp-mode nccl
