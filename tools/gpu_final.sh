#!/bin/bash
# One-GPU round-end check: GPU test suite, smoke, default bench line (+ reference arm), ncu launch list of one bench step and
# `ncu --set full` captures of the kernels DESIGN.md quotes.   usage (repo root, on the GPU box): bash tools/gpu_final.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/final.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/final.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | cut -c1-1200 | tee -a gpurun_out/final.log; }
TO=900 TAILN=8 run ${TAG}_final_tests python -m pytest tests -q -m gpu --no-header -p no:cacheprovider
TO=200 TAILN=5 run ${TAG}_final_smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TO=600 TAILN=3 run ${TAG}_final_bench python bench.py
TO=300 TAILN=3 run ${TAG}_final_bench_ref python bench.py --impl reference --steps 1 --warmup 0
TO=300 TAILN=4 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:k_attn_fwd -s 1 -c 1 -o gpurun_out/${TAG}_attn -f python tools/ncu_target.py attn
TO=300 TAILN=4 run ncu_attn_cp8 ncu --set full --clock-control none --import-source on -k regex:k_attn_fwd -s 1 -c 1 -o gpurun_out/${TAG}_attn_cp8 -f python tools/ncu_target.py attn_cp8
TO=300 TAILN=4 run ncu_gemm ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 2 -c 2 -o gpurun_out/${TAG}_gemm -f python tools/ncu_target.py gemm
TO=300 TAILN=4 run ncu_warp ncu --set full --clock-control none --import-source on -k regex:"k_splat|k_normalise|k_project" -s 4 -c 3 -o gpurun_out/${TAG}_splat -f python tools/ncu_target.py warp
TO=120 TAILN=8 run ${TAG}_attn_trace_1t python tools/attn_trace1t.py
# launch list of the whole bench process, Path D step + extras + Path R leg (cold-cache, serialised: compare SHARES)
TO=600 TAILN=3 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras
