#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 1200 gpurun_out/bench_final.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final_fp16x2.csv python tools/ncu_target.py fp16x2 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final_fp16x3.csv python tools/ncu_target.py fp16x3 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final_fp16.csv python tools/ncu_target.py fp16 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_layer -s 1 -c 1 -o gpurun_out/prof_stack_fp16x2 -f python tools/ncu_target.py fp16x2 3 > gpurun_out/ncu_stack2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_layer -s 1 -c 1 -o gpurun_out/prof_stack_fp16x3 -f python tools/ncu_target.py fp16x3 3 > gpurun_out/ncu_stack.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_layer -s 1 -c 1 -o gpurun_out/prof_stack_fp16 -f python tools/ncu_target.py fp16 3 > gpurun_out/ncu_stack16.log 2>&1
ls -la gpurun_out/*.ncu-rep
