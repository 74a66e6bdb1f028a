#!/bin/bash
# round-2 GPU session 31: two channel tiles per CTA in the transposed dense kernel
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_mlp_gpu.py tests/test_tc_gpu.py tests/test_models_gpu.py -q -x > gpurun_out/r02_t31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t31.log
tail -4 gpurun_out/r02_t31.log
timeout -k 10 900 python bench.py --steps 100 --warmup 8 --no-train > gpurun_out/r02_bench_cp2.json 2> gpurun_out/r02_bench_cp2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_cp2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['one_step_at_a_time']['ms_per_step'], d['e2e']['value'])
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
PY
