#!/bin/bash
# round-2 GPU session 16/17: kNN ordering by fine distances, canonical evaluation of the ambiguous entries only (17: compact ambiguous list)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -k "knn or dgcnn" > gpurun_out/r02_t17_knn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t17_knn.log
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_errstat.so timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_errstat2.json 2>gpurun_out/r02_knn_errstat2.err
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag8.json 2>gpurun_out/r02_knn_diag8.err
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_knn_full8 -f python tools/knn_tc_timing.py > gpurun_out/r02_ncu_knn8.log 2>&1
tail -3 gpurun_out/r02_t17_knn.log; cat gpurun_out/r02_knn_errstat2.json; cat gpurun_out/r02_knn_diag8.json
