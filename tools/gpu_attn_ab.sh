#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/attn_ab.log
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r02_tests_1t.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r02_tests_1t.log
for v in "G3C_ATTN_SHORT1T=0" "G3C_ATTN_SHORT1T=1"; do
  echo "== $v" | tee -a gpurun_out/attn_ab.log
  env $v G3C_PERF_LOG2=1 timeout 200 python tools/gpu_perf.py attn 2>&1 | grep '"attn"' | cut -c1-300 | tee -a gpurun_out/attn_ab.log
done
G3C_ATTN_SHORT1T=1 timeout 300 python -m pytest tests/test_dit_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "attention" 2>&1 | tail -3 | tee -a gpurun_out/attn_ab.log
