#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary4.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary4.txt; tail -n 14 gpurun_out/$name.log | cut -c1-1800 | tee -a gpurun_out/summary4.txt; }
run t_tc python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -s -x
run bench python bench.py --steps 3 --warmup 3
run trace32 python tools/trace_layer.py fp16x2 2
