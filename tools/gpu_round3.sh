#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary3.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary3.txt; tail -n 14 gpurun_out/$name.log | cut -c1-1500 | tee -a gpurun_out/summary3.txt; }
run t_tc python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "not fp32 or full_size" -s -x
run trace16 python tools/trace_layer.py fp16 2
run trace48 python tools/trace_layer.py fp16x3 2
run timing python tools/quick_timing.py
