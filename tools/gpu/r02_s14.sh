#!/bin/bash
# round-2 GPU session 14: ordered scatter-add gradients + whole GPU suite
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_t14_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t14_all.log
tail -5 gpurun_out/r02_t14_all.log
