#!/bin/bash
# round-2 GPU session 8: kNN tensor-core v2 (parity, phases, ncu), F1 v3e
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -q -k "knn_graph_tensor_core or dgcnn_graph" > gpurun_out/r02_t8_knn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t8_knn.log
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag2.json 2>gpurun_out/r02_knn_diag2.err
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_knn_launches2.csv python tools/knn_tc_timing.py > /dev/null 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_knn_full -f python tools/knn_tc_timing.py > gpurun_out/r02_ncu_knn.log 2>&1
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py -x -q -k "conv1_prebn" > gpurun_out/r02_t8_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t8_f1.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3e.json 2>gpurun_out/r02_f1v3e.err
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3e_timeline.json 2>&1
tail -3 gpurun_out/r02_t8_knn.log; cat gpurun_out/r02_knn_diag2.json; tail -3 gpurun_out/r02_t8_f1.log; cat gpurun_out/r02_f1v3e*.json
python - <<'PY'
import csv, re, collections
lines = open("gpurun_out/r02_knn_launches2.csv").readlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
agg = collections.defaultdict(list)
for r in csv.DictReader(lines[start:]):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        agg[re.sub(r"\(.*", "", r["Kernel Name"])[:60]].append(float(r["Metric Value"].replace(",", "")) / 1000.0)
for k, v in agg.items():
    if "knn" in k: print(k, [round(x, 1) for x in v[:12]])
PY
