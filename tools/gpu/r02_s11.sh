#!/bin/bash
# round-2 GPU session 11: synchronous F1 kernel (parity, timing, timeline) vs the producer/consumer variant
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py tests/test_train_gpu.py -x -q -k "conv1_prebn or training_step or first_layer" > gpurun_out/r02_t11_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t11_f1.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v4.json 2>gpurun_out/r02_f1v4.err
PSA_F1_VARIANT=2 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v4_pc.json 2>&1
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v4_timeline.json 2>&1
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1v4_sa2.json 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:sa_conv1_sync --launch-skip 1 -c 1 -o gpurun_out/r02_f1v4_full -f python tools/profile_ops.py > gpurun_out/r02_ncu_f1v4.log 2>&1
tail -3 gpurun_out/r02_t11_f1.log; cat gpurun_out/r02_f1v4*.json; tail -2 gpurun_out/r02_f1v4.err
