#!/bin/bash
# round-2 GPU session 15 (two GPUs): the N=2 bench exactly as the driver launches it, the reference arm, the NCCL training test
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_gpus.txt
timeout -k 10 600 python -m pytest tests/test_train_multi_gpu.py -q -x > gpurun_out/r02_t15_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t15_multi.log
tail -3 gpurun_out/r02_t15_multi.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"
tail -c 3000 gpurun_out/r02_bench_n2.json
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; echo "ref n2 rc=$?"
tail -c 1500 gpurun_out/r02_bench_ref_n2.json
tail -5 gpurun_out/r02_bench_n2.err
