#!/bin/bash
# round-2 GPU session 51: phase accounting of the 12-warp F1 kernel
mkdir -p gpurun_out
PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py 2>&1 | tail -1
