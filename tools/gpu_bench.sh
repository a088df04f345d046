#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; tail -c 3000 gpurun_out/bench_main.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 800 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_fp16x3.csv python tools/ncu_target.py fp16x3 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_fp16.csv python tools/ncu_target.py fp16 3 > /dev/null 2>&1
