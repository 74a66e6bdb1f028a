#!/bin/bash
# round-2 GPU session 45: F1 per-phase cycle accounting (producer: wait-empty / search / barrier / extract+rows; consumer: wait-full / conv+store)
mkdir -p gpurun_out
PSA_F1_VARIANT=4 PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_s45_static_tlog.json 2>&1; cat gpurun_out/r02_f1_s45_static_tlog.json
PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_s45_dyn_tlog.json 2>&1; cat gpurun_out/r02_f1_s45_dyn_tlog.json
