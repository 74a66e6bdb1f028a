#!/bin/bash
# round-2 GPU session 50: F1 per-CTA timeline with SM ids (which CTAs finish last?)
mkdir -p gpurun_out
PSA_F1_TLOG_DUMP=gpurun_out/r02_f1_s50_tlog.npy PSA_LIB_PATH=scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py 2>&1 | tail -1
