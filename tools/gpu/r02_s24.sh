#!/bin/bash
# round-2 GPU session 24: what the driver runs at round end -- smoke, the default bench line (timed), the reference arm
mkdir -p gpurun_out
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
t0=$(date +%s)
timeout -k 10 1500 python bench.py > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout -k 10 900 python bench.py --impl reference > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/r02_bench_final_reference_arm.err; echo "ref rc=$? wall=$(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('one_step_at_a_time'), d['e2e'], d['gpu_launches'])
print(d['roofline']); print(d['roofline_f1']['frac'], d['roofline_f1']['steady_state']['frac'], d['roofline_f1']['traffic'])
print(d['train']); print(d['cpu_baseline']); print(d['clocks'])
print({k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items() if isinstance(v,dict)})
r=json.loads(open('gpurun_out/r02_bench_final_reference_arm.json').read().strip().splitlines()[-1])
print(r['value'], r['cpu_baseline'])
PY
