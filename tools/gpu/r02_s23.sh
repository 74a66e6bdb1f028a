#!/bin/bash
# round-2 GPU session 23: full suite on the current build, bench, ncu --set full of the SA2 level
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r02_t23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t23.log
tail -4 gpurun_out/r02_t23.log
timeout -k 10 900 python bench.py --steps 40 --warmup 5 --no-train --no-extra > gpurun_out/r02_bench_f16c.json 2> gpurun_out/r02_bench_f16c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_f16c.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('one_step_at_a_time'), d['e2e'])
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in d.get('kernels',{}).items()})
PY
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_sa_dual_kernel --launch-skip 6 -c 1 -o gpurun_out/r02_dual_sa2_full -f python tools/profile_step.py 3 > gpurun_out/r02_ncu_dual_sa2.log 2>&1
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag11.json 2>gpurun_out/r02_knn_diag11.err; cat gpurun_out/r02_knn_diag11.json
