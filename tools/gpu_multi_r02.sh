#!/bin/bash
# Round-2 multi-GPU check (gpurun --gpus 4 -- 'bash tools/gpu_multi_r02.sh'): sharded-step parity tests, then the bench in
# the default CFG x CP layout at N = 2 and 4 and in the reference's cp-only layout at N = 4.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 500 python -m pytest tests/test_cp_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02_cp_pytest.log
tail -8 gpurun_out/r02_cp_pytest.log
run() { n=$1; tag=$2; shift 2; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 3 --warmup 3 "$@" > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err; echo "$tag rc=$?"; grep '"metric"' gpurun_out/r02_bench_$tag.json | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln); print(d['n_gpus'], d['config']['parallelism'], 'steps/s', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'attn TF/s', round(d['roofline']['achieved'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_breakdown'].items()}, d.get('sharded_parity'), d['clocks'])
"; tail -2 gpurun_out/r02_bench_$tag.err | cut -c1-300; }
run 2 n2
if [ "$N" -ge 4 ]; then run 4 n4; run 4 n4_cp --parallelism cp; fi
if [ "$N" -ge 8 ]; then run 8 n8; run 8 n8_cp --parallelism cp; fi
