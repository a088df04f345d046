#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep.py > gpurun_out/sweep.log 2>&1; tail -20 gpurun_out/sweep.log | cut -c1-260
