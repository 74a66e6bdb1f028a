#!/bin/bash
# round-2 GPU session 39: evidence for HEAD (transposed last layer in SA1): full GPU suite, ncu of the SA levels, launch list, default bench + reference arm
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_t39.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t39.log; tail -3 gpurun_out/r02_t39.log
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:tc_sa_dual_kernel --launch-skip 4 -c 4 -o gpurun_out/r02_dual_s39 -f python tools/profile_step.py 3 > gpurun_out/r02_ncu_dual_s39.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches_s39.csv python bench.py --steps 2 --warmup 1 --no-extra --no-train --no-cpu-baseline > gpurun_out/r02_bench_under_ncu_s39.log 2>&1
( time timeout -k 10 1200 python bench.py > gpurun_out/r02_bench_s39_1gpu.json 2> gpurun_out/r02_bench_s39_1gpu.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_s39_reference_arm.json 2> gpurun_out/r02_bench_s39_reference_arm.err ) 2>&1 | grep real; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_s39_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['one_step_at_a_time']['ms_per_step'], d['e2e']['value'], d['train']['ms_per_step'], d['roofline']['frac'], d['roofline']['tensor_pipe_frac'], d['roofline_f1']['frac'], d['gpu_launches'])
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
PY
