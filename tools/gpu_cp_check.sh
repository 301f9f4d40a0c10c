#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe_cpe.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe_cpe.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe_cpe.log; }
TO=200 TAILN=10 run cp_test5 python -m pytest tests/test_cp_gpu.py -q --no-header -p no:cacheprovider -s
TO=200 TAILN=3 run bench_n2_ce2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --cp-mode p2p --no-cpu-baseline
TO=200 TAILN=3 run bench_n2_nccl2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --cp-mode nccl --no-cpu-baseline
