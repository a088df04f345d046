#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "exit $?"; wc -l gpurun_out/bench_check.json; python -c "
import json; d=json.load(open('gpurun_out/bench_check.json')); print(d['value'], d['roofline']['traffic'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"; tail -3 gpurun_out/bench_check.err
