#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_cp_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r02_cp_pytest_2gpu.log
tail -6 gpurun_out/r02_cp_pytest_2gpu.log
run() { n=$1; tag=$2; shift 2; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 3 --warmup 3 --no-extras --no-path-r "$@" > gpurun_out/r02b_bench_$tag.json 2> gpurun_out/r02b_bench_$tag.err; echo "$tag rc=$?"; grep '"metric"' gpurun_out/r02b_bench_$tag.json | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln); print(d['n_gpus'], d['config']['parallelism'], 'steps/s', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), {k:round(v['ms_per_step'],1) for k,v in d['kernel_breakdown'].items()}, d.get('sharded_parity'), d['clocks'])
"; tail -2 gpurun_out/r02b_bench_$tag.err | cut -c1-300; }
run 2 n2
run 2 n2_cp --parallelism cp
