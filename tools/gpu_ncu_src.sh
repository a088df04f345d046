#!/bin/bash
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_tc_layer -s 45 -c 1 -o gpurun_out/prof_layer_v4_fp16 -f python tools/ncu_target.py fp16 4 > gpurun_out/ncu_v4.log 2>&1
tail -2 gpurun_out/ncu_v4.log
