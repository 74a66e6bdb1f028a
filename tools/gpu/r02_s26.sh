#!/bin/bash
# round-2 GPU session 26: training tests incl. mlp_training / FP module / pointnet2_cls_bga training
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_train_gpu.py -q -x > gpurun_out/r02_t26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t26.log
tail -30 gpurun_out/r02_t26.log
