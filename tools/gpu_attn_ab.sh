#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/attn_ab_step.log
timeout 300 python -m pytest tests/test_dit_ops_gpu.py tests/test_fullsize_properties_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "attention" 2>&1 | tail -3 | tee -a gpurun_out/attn_ab_step.log
for v in "G3C_ATTN_QT=0"; do
  echo "== $v" | tee -a gpurun_out/attn_ab_step.log
  env $v timeout 300 python bench.py --no-cpu-baseline --no-extras --no-path-r 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_breakdown'].items() if isinstance(v,dict)}, d['clocks'])" | tee -a gpurun_out/attn_ab_step.log
done
timeout 100 python tools/attn_trace1t.py 2>&1 | tail -8 | tee -a gpurun_out/attn_ab_step.log
