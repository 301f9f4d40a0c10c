#!/bin/bash
# smallest end-to-end check of the built library on a GPU box: engine tests + smoke
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_dit_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/quick.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/quick.log
