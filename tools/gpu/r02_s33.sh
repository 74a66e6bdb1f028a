#!/bin/bash
# round-2 GPU session 33 (two GPUs): final default bench lines at N=1 (one rank) and N=2 (torchrun), reference arm
mkdir -p gpurun_out
timeout -k 10 1500 python bench.py > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$?"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 > gpurun_out/r02_bench_final_2gpu.json 2> gpurun_out/r02_bench_final_2gpu.err; echo "bench n2 rc=$?"
timeout -k 10 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/r02_bench_final_reference_arm.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_final_1gpu.json','gpurun_out/r02_bench_final_2gpu.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('one_step_at_a_time') or {}).get('ms_per_step'), (d['train'] or {}).get('ms_per_step'), (d['train'] or {}).get('allreduce_us'))
    print('   ', {k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
r=json.loads(open('gpurun_out/r02_bench_final_reference_arm.json').read().strip().splitlines()[-1])
print(r['value'], r['cpu_baseline']['sample'])
PY
