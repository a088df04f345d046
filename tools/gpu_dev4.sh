#!/bin/bash
# development GPU call: GPU parity suite, config-4 (PNDM) bench line + ncu launch list, small-batch timeline
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -x -m gpu --timeout 300 > gpurun_out/dev_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/dev_tests.log
timeout 300 python bench.py --config 4 --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/dev_bench4.json 2> gpurun_out/dev_bench4.err; echo "bench4 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/dev_bench4.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['avg_launch_us'],d['roofline']['stack_only']['frac'],d['roofline']['stack_only']['avg_launch_us'],d['e2e']['value'],d['clocks'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/dev_launches_config4.csv python tools/ncu_target.py fp16s 3 4 > /dev/null 2>&1
grep k_tc_stack gpurun_out/dev_launches_config4.csv | tail -6 | cut -d, -f5,15- | cut -c1-120
timeout 200 python tools/dev_stack.py trace 1 512 > gpurun_out/dev_trace_1x512.log 2>&1; echo "trace rc=$?"; grep -E "^--- fp16s|cycles per layer|fused head|end of skip|producer 0 \(" gpurun_out/dev_trace_1x512.log | head -6
