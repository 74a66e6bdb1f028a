#!/bin/bash
# round-2 GPU session 32: batches in flight (engine slots) 4 vs 6 vs 8, then the final default bench line + reference arm
mkdir -p gpurun_out
for s in 6 8; do
  timeout -k 10 600 python bench.py --steps 200 --warmup 8 --streams $s --no-train --no-extra --no-cpu-baseline > gpurun_out/r02_bench_streams$s.json 2> gpurun_out/r02_bench_streams$s.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_streams$s.json').read().strip().splitlines()[-1])
print('streams=$s', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
PY
done
timeout -k 10 1500 python bench.py > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['one_step_at_a_time']['ms_per_step'], d['e2e']['value'], d['train']['ms_per_step'], d['roofline']['frac'], d['roofline']['tensor_pipe_frac'], d['roofline_f1']['frac'])
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
PY
