#!/bin/bash
# round-2 GPU session 44: F1 with dynamic unit distribution inside a cloud (A/B against static ranges), parity, timeline, ncu
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py tests/test_train_gpu.py tests/test_edge_cases_gpu.py -x -q -k "conv1_prebn or train or sa_module" > gpurun_out/r02_t44.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t44.log; tail -4 gpurun_out/r02_t44.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_s44.json 2> gpurun_out/r02_f1_s44.err; cat gpurun_out/r02_f1_s44.json
PSA_F1_VARIANT=4 timeout -k 10 120 python tools/f1_timing.py 2>&1 | tail -1
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1_s44_sa2.json 2>&1; cat gpurun_out/r02_f1_s44_sa2.json
PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_s44_tlog.json 2>&1; cat gpurun_out/r02_f1_s44_tlog.json
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:sa_conv1_stream_kernel --launch-skip 3 -c 1 -o gpurun_out/r02_f1_s44 -f python tools/f1_timing.py > gpurun_out/r02_ncu_f1_s44.log 2>&1; tail -2 gpurun_out/r02_ncu_f1_s44.log
