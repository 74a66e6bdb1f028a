#!/bin/bash
# round-2 GPU session 19: fp16x2 operand split with the bf16x3 range guard -- parity tests, then the bench
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_mlp_gpu.py tests/test_tc_gpu.py tests/test_models_gpu.py tests/test_train_gpu.py -q -x > gpurun_out/r02_t19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t19.log
tail -15 gpurun_out/r02_t19.log
timeout -k 10 900 python bench.py --steps 40 --warmup 5 --no-train > gpurun_out/r02_bench_f16.json 2> gpurun_out/r02_bench_f16.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_f16.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('one_step_at_a_time'), d['e2e'])
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
PY
