#!/bin/bash
# round-2 GPU session 53: F1 three-stage pipeline for levels without features, two-stage (search+extract | 8 conv warps) for levels with features
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py tests/test_train_gpu.py tests/test_edge_cases_gpu.py -x -q -k "conv1_prebn or train or sa_module" > gpurun_out/r02_t53.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t53.log; tail -4 gpurun_out/r02_t53.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_s53.json 2> gpurun_out/r02_f1_s53.err; cat gpurun_out/r02_f1_s53.json
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1_s53_sa2.json 2>&1; cat gpurun_out/r02_f1_s53_sa2.json
PSA_F1_TLOG_DUMP=gpurun_out/r02_f1_s53_tlog.npy PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_s53_tlog.json 2>&1; cat gpurun_out/r02_f1_s53_tlog.json
