#!/bin/bash
# First GPU pass: self-tests, parity tests split by path (a faulting kernel poisons its process only), quick timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n 25 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run selftest python -c "import diffsinger_b200 as d; rc, rep = d.selftest(0); print(rep); print('rc', rc)"
run t_fp32 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "fp32 and not full_size" -s
run t_tc_fwd python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "forward_golden and not fp32" -s
run t_tc_rest python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -k "not forward_golden and not fp32 or full_size" -s
run smoke python __graft_entry__.py --smoke
run timing python tools/quick_timing.py
