#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/attn_ab_step.log
for v in "G3C_ATTN_POLY1T=3" "G3C_ATTN_POLY1T=4"; do
  echo "== $v" | tee -a gpurun_out/attn_ab_step.log
  env $v timeout 300 python bench.py --no-cpu-baseline --no-extras --no-path-r 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_breakdown'].items() if isinstance(v,dict)}, d['clocks'])" | tee -a gpurun_out/attn_ab_step.log
done
