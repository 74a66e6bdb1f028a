#!/bin/bash
# round-2 GPU session 54 (last ~4 GPU-minutes): final-tree evidence in priority order, every step writes gpurun_out/ as it goes
mkdir -p gpurun_out
timeout -k 5 70 python tools/f1_timing.py > gpurun_out/r02_f1_s54.json 2> gpurun_out/r02_f1_s54.err; cat gpurun_out/r02_f1_s54.json
timeout -k 5 90 ncu --set full --clock-control none --import-source on -k regex:sa_conv1_stream_kernel --launch-skip 3 -c 1 -o gpurun_out/r02_f1_s54 -f python tools/f1_timing.py > gpurun_out/r02_ncu_f1_s54.log 2>&1; tail -2 gpurun_out/r02_ncu_f1_s54.log
timeout -k 5 120 python bench.py --steps 200 --warmup 8 --no-train --no-cpu-baseline --no-extra > gpurun_out/r02_bench_s54.json 2> gpurun_out/r02_bench_s54.err; tail -c 1500 gpurun_out/r02_bench_s54.json
timeout -k 5 120 ncu --set full --clock-control none --import-source on -k regex:tc_sa_dual_kernel --launch-skip 4 -c 2 -o gpurun_out/r02_dual_s54 -f python tools/profile_step.py 3 > gpurun_out/r02_ncu_dual_s54.log 2>&1; tail -2 gpurun_out/r02_ncu_dual_s54.log
