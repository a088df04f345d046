#!/bin/bash
# one full ncu capture (with source-level sampling) of a steady-state residual-stack launch + the launch list of the loop
mkdir -p gpurun_out
P=${1:-fp16s}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$P.csv python tools/ncu_target.py $P 3 > gpurun_out/ncu_list_$P.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tc_stack -s 1 -c 1 -f -o gpurun_out/ncu_src_$P python tools/ncu_target.py $P 3 > gpurun_out/ncu_src_$P.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_src_$P.log; ls -la gpurun_out/ncu_src_$P.ncu-rep
