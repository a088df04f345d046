#!/bin/bash
# 2-GPU bench line (weak-scaled config 2 + strong-scaled config 4 as `extra`), launched the way the driver does
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "rc=$?"
tail -c 1800 gpurun_out/r02_bench_2gpu.json; tail -3 gpurun_out/r02_bench_2gpu.err
