#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary6.txt; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary6.txt; tail -n 16 gpurun_out/$name.log | cut -c1-1200 | tee -a gpurun_out/summary6.txt; }
run t_tc python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "not fp32 or full_size"
run probe python tools/probe_layers.py
run trace32 python tools/trace_layer.py fp16x2 2
