#!/bin/bash
# round-2 GPU session 47: cycles per query of the F1 search inner loop, packed f32x2 vs scalar (tools/microbench/search_rate.cu)
mkdir -p gpurun_out
(cd tools/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/search_rate search_rate.cu && /tmp/search_rate) > gpurun_out/r02_search_rate.jsonl 2>&1
cat gpurun_out/r02_search_rate.jsonl
