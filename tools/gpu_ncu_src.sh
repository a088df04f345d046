#!/bin/bash
# one full ncu capture (with source-level sampling) of a steady-state residual-stack launch
mkdir -p gpurun_out
P=${1:-fp16x2}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tc_layer -s 1 -c 1 -f -o gpurun_out/ncu_src_$P python tools/ncu_target.py $P 3 > gpurun_out/ncu_src_$P.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_src_$P.log; ls -la gpurun_out/ncu_src_$P.ncu-rep
