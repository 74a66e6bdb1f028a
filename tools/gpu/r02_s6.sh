#!/bin/bash
# round-2 GPU session 6: tensor-core kNN graph (parity + timing), DGCNN at n=2048, F1 v3d (prefetched centres, register extraction)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -q -s -k "knn_graph_tensor_core or dgcnn_graph" > gpurun_out/r02_t6_knn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t6_knn.log
timeout -k 10 900 python -m pytest tests/test_models_gpu.py -q -k "dgcnn" > gpurun_out/r02_t6_dgcnn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t6_dgcnn.log
timeout -k 10 600 python -m pytest tests/test_mlp_gpu.py tests/test_train_gpu.py -x -q -k "conv1_prebn or training_step or first_layer" > gpurun_out/r02_t6_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t6_f1.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3d.json 2>gpurun_out/r02_f1v3d.err
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3d_timeline.json 2>&1
timeout -k 10 120 python tools/f1_timing.py sa2 > gpurun_out/r02_f1v3d_sa2.json 2>&1
timeout -k 10 300 python - > gpurun_out/r02_knn_timing.json 2>gpurun_out/r02_knn_timing.err <<'PY'
import json, torch
from scanobjectnn_b200 import ops
from scanobjectnn_b200.synthetic import make_clouds
B, N = 32, 2048
out = {}
flush = torch.zeros(64 * 1024 * 1024, device="cuda"); sink = torch.zeros((), device="cuda")
for name, x in (("c64", torch.randn((B, N, 64), device="cuda")), ("c3", torch.from_numpy(make_clouds("ball", B, N, seed=5)).cuda())):
    for mode in ("tc", "fp32"):
        ops._KNN_FP32_ONLY = mode == "fp32"
        for _ in range(3): ops.knn_graph(x, 20)
        ts = []
        for _ in range(7):
            sink.copy_(flush.sum()); torch.cuda._sleep(400_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = ops.knn_graph(x, 20); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort(); out[f"knn_graph_{name}_{mode}_us"] = ts[len(ts) // 2]
    ops._KNN_FP32_ONLY = False
print(json.dumps(out))
PY
tail -8 gpurun_out/r02_t6_knn.log; tail -3 gpurun_out/r02_t6_dgcnn.log; tail -3 gpurun_out/r02_t6_f1.log; cat gpurun_out/r02_f1v3d*.json gpurun_out/r02_knn_timing.json; tail -3 gpurun_out/r02_knn_timing.err
