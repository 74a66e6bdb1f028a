#!/bin/bash
# round-2 GPU session 3: training parity (full file), streaming F1 v3 (exhaustive register search), training launch list
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_train_gpu.py -q -s > gpurun_out/r02_t3_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t3_train.log
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py -x -q -k "conv1_prebn or unit_scale" > gpurun_out/r02_t3_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t3_f1.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3.json 2>gpurun_out/r02_f1v3.err
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3_timeline.json 2>&1
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1v3_sa2.json 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches.csv python tools/profile_train.py 2 > gpurun_out/r02_train_prof.log 2>&1
tail -25 gpurun_out/r02_t3_train.log; tail -4 gpurun_out/r02_t3_f1.log; cat gpurun_out/r02_f1v3*.json; tail -3 gpurun_out/r02_train_prof.log
