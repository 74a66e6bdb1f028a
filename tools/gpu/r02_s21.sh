#!/bin/bash
# round-2 GPU session 21: centred kNN passes, multi-layer EdgeConv on the set-abstraction kernel, wide D double-buffering, image twins
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r02_t21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t21.log
tail -8 gpurun_out/r02_t21.log
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag10.json 2>gpurun_out/r02_knn_diag10.err; cat gpurun_out/r02_knn_diag10.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_dgcnn_launches2.csv python tools/profile_dgcnn.py 2 > gpurun_out/r02_dgcnn_prof2.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_step_launches2.csv python tools/profile_step.py 2 > gpurun_out/r02_step_prof2.log 2>&1
timeout -k 10 900 python bench.py --steps 40 --warmup 5 --no-train > gpurun_out/r02_bench_f16b.json 2> gpurun_out/r02_bench_f16b.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_f16b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('one_step_at_a_time'), d['e2e'])
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
PY
