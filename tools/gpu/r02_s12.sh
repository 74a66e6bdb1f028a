#!/bin/bash
# round-2 GPU session 12: producer/consumer F1 at 3 CTAs per SM; kNN tensor-core v5 (running cut)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py -x -q -k "conv1_prebn" > gpurun_out/r02_t12_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t12_f1.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v5.json 2>gpurun_out/r02_f1v5.err
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v5_timeline.json 2>&1
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1v5_sa2.json 2>&1
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -q -k "knn_graph_tensor_core or dgcnn_graph" > gpurun_out/r02_t12_knn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t12_knn.log
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag5.json 2>gpurun_out/r02_knn_diag5.err
tail -3 gpurun_out/r02_t12_f1.log; cat gpurun_out/r02_f1v5*.json; tail -3 gpurun_out/r02_t12_knn.log; cat gpurun_out/r02_knn_diag5.json
