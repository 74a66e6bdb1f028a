#!/bin/bash
# round-2 GPU session 36: final evidence -- ncu --set full of both SA levels, launch list of the bench command, GPU suite, default bench + reference arm
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -q -m gpu > gpurun_out/r02_t36.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t36.log; tail -3 gpurun_out/r02_t36.log
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_sa_dual_kernel --launch-skip 4 -c 4 -o gpurun_out/r02_dual_final -f python tools/profile_step.py 3 > gpurun_out/r02_ncu_dual_final.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_dense3_kernel --launch-skip 14 -c 7 -o gpurun_out/r02_dense3_final -f python tools/profile_step.py 3 > gpurun_out/r02_ncu_dense3_final.log 2>&1
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-train --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout -k 10 1500 python bench.py > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$?"
timeout -k 10 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/r02_bench_final_reference_arm.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['one_step_at_a_time']['ms_per_step'], d['e2e']['value'], d['train']['ms_per_step'], d['roofline']['frac'], d['roofline']['tensor_pipe_frac'], d['roofline_f1']['frac'], d['gpu_launches'])
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
PY
