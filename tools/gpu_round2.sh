#!/bin/bash
# bench line + ncu launch list + one full ncu capture of the layer kernel
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary2.txt; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary2.txt; tail -n 12 gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary2.txt; }
run bench python bench.py --steps 3 --warmup 3
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fp16x3.csv python tools/ncu_target.py fp16x3 2
run ncu_list16 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fp16.csv python tools/ncu_target.py fp16 2
run ncu_full ncu --set full --clock-control none --import-source on -k regex:k_tc_layer -s 22 -c 2 -o gpurun_out/prof_layer_fp16x3 -f python tools/ncu_target.py fp16x3 2
run ncu_full16 ncu --set full --clock-control none --import-source on -k regex:k_tc_layer -s 22 -c 2 -o gpurun_out/prof_layer_fp16 -f python tools/ncu_target.py fp16 2
