#!/bin/bash
# round-2 GPU session 29: final state -- whole GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -q -m gpu > gpurun_out/r02_t29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t29.log
tail -4 gpurun_out/r02_t29.log
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 10 1500 python bench.py > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['one_step_at_a_time']['ms_per_step'], d['e2e']['value'], d['train']['ms_per_step'], d['roofline']['frac'], d['roofline_f1']['frac'])
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
PY
