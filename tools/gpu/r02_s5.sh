#!/bin/bash
# round-2 GPU session 5: training parity, F1 v3c (whole-row stores), ncu of the F1 kernel
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_train_gpu.py -q -s > gpurun_out/r02_t5_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t5_train.log
timeout -k 10 900 python -m pytest tests/test_mlp_gpu.py -x -q -k "conv1_prebn" > gpurun_out/r02_t5_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t5_f1.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3c.json 2>gpurun_out/r02_f1v3c.err
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3c_timeline.json 2>&1
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1v3c_sa2.json 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:sa_conv1_stream --launch-skip 1 -c 1 -o gpurun_out/r02_f1_full -f python tools/profile_ops.py > gpurun_out/r02_ncu_f1.log 2>&1
tail -6 gpurun_out/r02_t5_train.log; tail -3 gpurun_out/r02_t5_f1.log; cat gpurun_out/r02_f1v3c*.json; tail -3 gpurun_out/r02_ncu_f1.log
