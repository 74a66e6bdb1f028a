#!/bin/bash
# round-2 GPU session 1: write-bandwidth ceiling, streaming F1 kernel parity + timing (new vs round-1 kernel)
mkdir -p gpurun_out
(cd tools/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/write_bw write_bw.cu && /tmp/write_bw) > gpurun_out/r02_write_bw.jsonl 2>&1
timeout -k 10 400 python -m pytest tests/test_mlp_gpu.py tests/test_ops_gpu.py tests/test_edge_cases_gpu.py -x -q -k "conv1_prebn or ball or query" > gpurun_out/r02_t1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_t1.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_new.json 2>gpurun_out/r02_f1_new.err
PSA_F1_VARIANT=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1_old.json 2>&1
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1_new_sa2.json 2>&1
PSA_F1_VARIANT=1 timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1_old_sa2.json 2>&1
tail -5 gpurun_out/r02_t1.log; cat gpurun_out/r02_f1_*.json gpurun_out/r02_write_bw.jsonl
