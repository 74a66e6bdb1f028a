#!/bin/bash
# round-2 GPU session 18: kNN -- hoisted norm loads (default build) and the three-term (bf16x2) pass 2 variant
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -k "knn or dgcnn" > gpurun_out/r02_t18_knn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t18_knn.log
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_knn3.so timeout -k 10 600 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -k "knn or dgcnn" > gpurun_out/r02_t18_knn3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t18_knn3.log
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag9.json 2>gpurun_out/r02_knn_diag9.err
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_knn3.so timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag9_terms3.json 2>gpurun_out/r02_knn_diag9_terms3.err
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_knn_full9 -f python tools/knn_tc_timing.py > gpurun_out/r02_ncu_knn9.log 2>&1
tail -2 gpurun_out/r02_t18_knn.log; tail -2 gpurun_out/r02_t18_knn3.log; cat gpurun_out/r02_knn_diag9.json; cat gpurun_out/r02_knn_diag9_terms3.json
