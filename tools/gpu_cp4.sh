#!/bin/bash
mkdir -p gpurun_out
for mode in p2p nccl; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 2 --warmup 3 --cp-mode $mode > gpurun_out/bench_n4_$mode.log 2>&1
echo "$mode rc=$?"; grep '"metric"' gpurun_out/bench_n4_$mode.log | cut -c1-420; tail -3 gpurun_out/bench_n4_$mode.log | cut -c1-300
done
