#!/bin/bash
# round-2 GPU session 52: F1 as a three-stage pipeline (search | extract + rows | conv + store, one warpgroup each, setmaxnreg): parity, timing, stage accounting
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py tests/test_train_gpu.py tests/test_edge_cases_gpu.py -x -q -k "conv1_prebn or train or sa_module" > gpurun_out/r02_t52.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t52.log; tail -4 gpurun_out/r02_t52.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_s52.json 2> gpurun_out/r02_f1_s52.err; cat gpurun_out/r02_f1_s52.json
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1_s52_sa2.json 2>&1; cat gpurun_out/r02_f1_s52_sa2.json
PSA_F1_TLOG_DUMP=gpurun_out/r02_f1_s52_tlog.npy PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_s52_tlog.json 2>&1; cat gpurun_out/r02_f1_s52_tlog.json
