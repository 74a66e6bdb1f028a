#!/bin/bash
# round-2 GPU session 10: the whole GPU suite on the pruned library, kNN tensor-core v4 (subsampled pass 1)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02_t10_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t10_all.log
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag4.json 2>gpurun_out/r02_knn_diag4.err
tail -8 gpurun_out/r02_t10_all.log; cat gpurun_out/r02_knn_diag4.json
