#!/bin/bash
# round-2 GPU session 9: kNN tensor-core v3 (two threads per row), mbarrier back-off A/B on the inference step
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -q -k "knn_graph_tensor_core or dgcnn_graph" > gpurun_out/r02_t9_knn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t9_knn.log
timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_diag3.json 2>gpurun_out/r02_knn_diag3.err
for v in default sleep32 sleep100; do
  if [ $v = default ]; then unset PSA_LIB_PATH; else export PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_$v.so; fi
  timeout -k 10 600 python bench.py --steps 300 --warmup 10 --no-extra --no-train --no-cpu-baseline > gpurun_out/r02_bench_$v.json 2>gpurun_out/r02_bench_$v.err
done
unset PSA_LIB_PATH
tail -3 gpurun_out/r02_t9_knn.log; cat gpurun_out/r02_knn_diag3.json
python - <<'PY'
import json
for v in ("default", "sleep32", "sleep100"):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_{v}.json"))
        print(v, round(d["value"]), d["ms_per_step"], {k: round(d["kernels"][k]["us"], 1) for k in ("sa1_mlp", "sa2_mlp", "sa3_mlp")}, d["one_step_at_a_time"]["ms_per_step"])
    except Exception as e:
        print(v, "failed", e)
PY
