#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name ===" | tee -a gpurun_out/probe9.log; timeout "$TO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc" | tee -a gpurun_out/probe9.log; tail -n "${TAILN:-25}" gpurun_out/$name.log | tee -a gpurun_out/probe9.log; }
TO=300 TAILN=80 run trace python tools/attn_trace.py
TO=600 TAILN=15 run alltests python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x
