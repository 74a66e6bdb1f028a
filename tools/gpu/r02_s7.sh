#!/bin/bash
# round-2 GPU session 7: kNN tensor-core diagnostics (phases, exhaustive-row count), tensor-core training forward parity + timing
mkdir -p gpurun_out
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag.json 2>gpurun_out/r02_knn_diag.err
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_knn_launches.csv python tools/knn_tc_timing.py > /dev/null 2>&1
timeout -k 10 900 python -m pytest tests/test_train_gpu.py -q -s > gpurun_out/r02_t7_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t7_train.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches_c.csv python tools/profile_train.py 2 > gpurun_out/r02_train_prof_c.log 2>&1
cat gpurun_out/r02_knn_diag.json; tail -3 gpurun_out/r02_knn_diag.err; tail -6 gpurun_out/r02_t7_train.log
python - <<'PY'
import csv, re, collections
for f in ("gpurun_out/r02_knn_launches.csv",):
    lines = open(f).readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    agg = collections.defaultdict(list)
    for r in csv.DictReader(lines[start:]):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            agg[re.sub(r"\(.*", "", r["Kernel Name"])[:60]].append(float(r["Metric Value"].replace(",", "")) / 1000.0)
    for k, v in agg.items():
        if "knn" in k: print(k, [round(x, 1) for x in v[:12]])
PY
