#!/bin/bash
# round-2 GPU session 13: kNN tensor-core with four threads per row
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -q -k "knn_graph_tensor_core or dgcnn_graph" > gpurun_out/r02_t13_knn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t13_knn.log
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag6.json 2>gpurun_out/r02_knn_diag6.err
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_knn_full6 -f python tools/knn_tc_timing.py > gpurun_out/r02_ncu_knn6.log 2>&1
tail -3 gpurun_out/r02_t13_knn.log; cat gpurun_out/r02_knn_diag6.json
