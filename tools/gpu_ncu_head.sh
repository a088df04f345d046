#!/bin/bash
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_tc_head -s 3 -c 1 -o gpurun_out/prof_head_fp16 -f python tools/ncu_target.py fp16 4 > gpurun_out/ncu_head.log 2>&1
tail -2 gpurun_out/ncu_head.log
