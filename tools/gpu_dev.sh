#!/bin/bash
# one development GPU call: first contact, timeline, the GPU parity suite, a short bench line
mkdir -p gpurun_out
timeout 200 python tools/dev_stack.py quick > gpurun_out/dev_quick.log 2>&1; echo "quick rc=$?"; tail -6 gpurun_out/dev_quick.log
timeout 200 python tools/dev_stack.py k100 2>&1 | grep "ddpm K"
timeout 200 python tools/dev_stack.py trace > gpurun_out/dev_trace.log 2>&1; echo "trace rc=$?"; grep -E "cycles per layer|CTAs|fused head|end of skip" gpurun_out/dev_trace.log
timeout 700 python -m pytest tests -q -x -m gpu --timeout 300 > gpurun_out/dev_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/dev_tests.log
