#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py fp16x2 fp16x3 > gpurun_out/sanitize.log 2>&1
echo "exit $?"; grep -c "Invalid\|error" gpurun_out/sanitize.log; tail -15 gpurun_out/sanitize.log | cut -c1-300
