#!/bin/bash
# round-2 GPU session 25 (two GPUs): training tests incl. the per-level autograd path, the two-rank NCCL test, the N=2 bench lines
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_train_gpu.py tests/test_train_multi_gpu.py -q -x > gpurun_out/r02_t25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t25.log
tail -6 gpurun_out/r02_t25.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r02_bench_final_2gpu.json 2> gpurun_out/r02_bench_final_2gpu.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_final_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['train'])
PY
tail -3 gpurun_out/r02_bench_final_2gpu.err
