#!/bin/bash
# round-end validation of the committed build: GPU parity suite, smoke, ncu launch list + one full capture of the step kernel, a bench line
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_fp16s.csv python tools/ncu_target.py fp16s 3 > /dev/null 2>&1; grep -c k_tc_stack gpurun_out/r02_launches_fp16s.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tc_stack -s 1 -c 1 -f -o gpurun_out/r02_ncu_stack_fp16s python tools/ncu_target.py fp16s 3 > gpurun_out/r02_ncu_stack_fp16s.log 2>&1; tail -2 gpurun_out/r02_ncu_stack_fp16s.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_config4_fp16s.csv python tools/ncu_target.py fp16s 3 4 > /dev/null 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['avg_launch_us'],d['roofline']['stack_only']['frac'],d['e2e']['value'],d['clocks'])"
