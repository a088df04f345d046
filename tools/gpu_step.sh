#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_breakdown.py fp16x2 fp16 > gpurun_out/step_breakdown.log 2>&1; tail -4 gpurun_out/step_breakdown.log
