#!/bin/bash
# usage: gpu_scale.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "rc=$?"; grep '"metric"' gpurun_out/bench_n$N.log | cut -c1-1500; tail -5 gpurun_out/bench_n$N.log | cut -c1-300
