#!/bin/bash
# round-2 GPU session 38: transposed last layer of 64-wide SA levels (tc_sa_dual_kernel<3,2>)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_mlp_gpu.py tests/test_tc_gpu.py tests/test_models_gpu.py -q -x > gpurun_out/r02_t38.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t38.log
tail -12 gpurun_out/r02_t38.log
for tm in 1 0; do
PSA_SA_TMODE=$tm timeout -k 10 900 python bench.py --steps 200 --warmup 8 --no-train --no-cpu-baseline > gpurun_out/r02_bench_tmode$tm.json 2> gpurun_out/r02_bench_tmode$tm.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_tmode$tm.json').read().strip().splitlines()[-1])
print('tmode=$tm', d['value'], d['ms_per_step'], d['one_step_at_a_time']['ms_per_step'], d['e2e']['value'])
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
PY
done
