#!/bin/bash
# round-2 GPU session 2: training-path parity, streaming F1 v2 (timing + phase timeline), quick bench line
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_train_gpu.py -x -q -s > gpurun_out/r02_t2_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t2_train.log
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py tests/test_engine_gpu.py tests/test_models_gpu.py -x -q -k "conv1_prebn or engine or bga" > gpurun_out/r02_t2_misc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t2_misc.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v2.json 2>gpurun_out/r02_f1v2.err
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v2_timeline.json 2>&1
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1v2_sa2.json 2>&1
timeout -k 10 900 python bench.py --steps 40 --warmup 3 --train-steps 5 > gpurun_out/r02_bench_quick.json 2>gpurun_out/r02_bench_quick.err; echo "bench rc=$?" >> gpurun_out/r02_bench_quick.err
tail -15 gpurun_out/r02_t2_train.log; tail -4 gpurun_out/r02_t2_misc.log; cat gpurun_out/r02_f1v2*.json; tail -5 gpurun_out/r02_bench_quick.err; head -c 3000 gpurun_out/r02_bench_quick.json
