#!/bin/bash
# A/B: k100 error + timeline + short bench (no tests)
mkdir -p gpurun_out
timeout 200 python tools/dev_stack.py k100 2>&1 | grep "ddpm K"
timeout 200 python tools/dev_stack.py trace > gpurun_out/dev_trace.log 2>&1; echo "trace rc=$?"; grep -E "cycles per layer|CTAs|fused head|end of skip" gpurun_out/dev_trace.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/dev_bench.json 2> gpurun_out/dev_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/dev_bench.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['avg_launch_us'],d['roofline']['stack_only']['frac'],d['roofline']['stack_only']['avg_launch_us'],d['e2e']['value'],d['clocks'])"
