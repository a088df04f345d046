#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary5.txt; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary5.txt; tail -n 40 gpurun_out/$name.log | cut -c1-900 | tee -a gpurun_out/summary5.txt; }
run probe python tools/probe_prefetch.py
run trace_p0 python tools/trace_layer.py fp16x2 2 0
run trace_p1 python tools/trace_layer.py fp16x2 2 1
run trace_p2 python tools/trace_layer.py fp16x2 2 2
