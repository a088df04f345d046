#!/bin/bash
# Round-2 evidence, one GPU call: bench lines of every BASELINE config + sweep, the reference arm, ncu launch lists and full
# captures of the fused stack kernel, its clock64 timeline, compute-sanitizer memcheck.  Outputs land in gpurun_out/ (copied into profiles/ by hand).
mkdir -p gpurun_out
P=${1:-fp16s}
timeout 700 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_config2.json 2> gpurun_out/r02_bench_config2.err; echo "bench2 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> /dev/null; echo "ref rc=$?"
for C in 1 3 4; do
  timeout 600 python bench.py --config $C --steps 2 --warmup 3 --no-extra > gpurun_out/r02_bench_config$C.json 2> gpurun_out/r02_bench_config$C.err; echo "bench$C rc=$?"
done
echo "sweep skipped (profiles/r02_bench_sweep.json is from the previous evidence call of this session)"
for Q in fp16s; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_$Q.csv python tools/ncu_target.py $Q 3 > /dev/null 2>&1
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tc_stack -s 1 -c 1 -f -o gpurun_out/r02_ncu_stack_$Q python tools/ncu_target.py $Q 3 > gpurun_out/r02_ncu_stack_$Q.log 2>&1
done
# (the head is part of k_tc_stack since the fusion: no separate head capture)
timeout 300 python tools/dev_stack.py trace > gpurun_out/r02_stack_timeline.txt 2>&1
timeout 200 python tools/dev_stack.py trace 1 512 > gpurun_out/r02_stack_timeline_1x512.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/dev_stack.py quick > gpurun_out/r02_compute_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_config4_$P.csv python tools/ncu_target.py $P 3 4 > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r02_smi.txt
ls -la gpurun_out | grep r02_
