#!/bin/bash
# Round-end refresh after an attention-only change: GPU suite, smoke, default bench line, ncu captures of the attention kernel,
# its clock64 trace and the launch list of one bench step (GEMM / splat captures of tools/gpu_final.sh stay valid).
TAG=${1:-r02}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/final.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/final.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | cut -c1-1500 | tee -a gpurun_out/final.log; }
TO=900 TAILN=6 run ${TAG}_final_tests python -m pytest tests -q -m gpu --no-header -p no:cacheprovider
TO=200 TAILN=3 run ${TAG}_final_smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TO=600 TAILN=3 run ${TAG}_final_bench python bench.py
TO=300 TAILN=2 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:k_attn_fwd -s 1 -c 1 -o gpurun_out/${TAG}_attn -f python tools/ncu_target.py attn
TO=300 TAILN=2 run ncu_attn_cp8 ncu --set full --clock-control none --import-source on -k regex:k_attn_fwd -s 1 -c 1 -o gpurun_out/${TAG}_attn_cp8 -f python tools/ncu_target.py attn_cp8
TO=120 TAILN=9 run ${TAG}_attn_trace_1t python tools/attn_trace1t.py
TO=600 TAILN=2 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras
