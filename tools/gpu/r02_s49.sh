#!/bin/bash
# round-2 GPU session 49: F1 A/B without the output stream (is the kernel bound by its stores?)
mkdir -p gpurun_out
PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1nostore.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py 2>&1 | tail -1
PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py 2>&1 | tail -1
