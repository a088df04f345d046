#!/bin/bash
# development GPU call: first contact, timeline, one full ncu capture (source-level) of a steady-state fused launch, short bench
mkdir -p gpurun_out
timeout 200 python tools/dev_stack.py quick > gpurun_out/dev_quick.log 2>&1; echo "quick rc=$?"; tail -5 gpurun_out/dev_quick.log
timeout 200 python tools/dev_stack.py trace > gpurun_out/dev_trace.log 2>&1; echo "trace rc=$?"; grep -E "cycles per layer|CTAs|fused head|end of skip|producer 0 \(" gpurun_out/dev_trace.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tc_stack -s 1 -c 1 -f -o gpurun_out/ncu_src_fp16s python tools/ncu_target.py fp16s 3 > gpurun_out/ncu_src_fp16s.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_src_fp16s.log; ls -la gpurun_out/ncu_src_fp16s.ncu-rep
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/dev_bench.json 2> gpurun_out/dev_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/dev_bench.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['avg_launch_us'],d['roofline']['stack_only']['frac'],d['roofline']['stack_only']['avg_launch_us'],d['e2e']['value'],d['clocks'])"
