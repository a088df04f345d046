#!/bin/bash
# ncu evidence for profiles/: launch lists of a 3-step loop per precision + one full capture of a steady-state stack launch
mkdir -p gpurun_out
for P in fp16x2 fp16x3 fp16; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file gpurun_out/launches_final_$P.csv python tools/ncu_target.py $P 3 > /dev/null 2>&1
done
for P in fp16x2 fp16x3 fp16; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tc_layer -s 1 -c 1 -o gpurun_out/prof_stack_$P -f python tools/ncu_target.py $P 3 > gpurun_out/ncu_stack_$P.log 2>&1
done
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_final_*.csv
