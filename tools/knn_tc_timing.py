"""Tensor-core kNN graph (csrc/knn_tc.cu) at B=32, N=2048, k=20: time per phase (prep / main / exhaustive rows) with CUDA events
between the launches (PSA_KNN_PHASES build not needed: the three kernels are separate launches), number of rows that went to the
exhaustive kernel, fp32 kernel beside it."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from scanobjectnn_b200 import _lib, ops
from scanobjectnn_b200.synthetic import make_clouds

lib = _lib.load()
B, N, K = 32, 2048, 20
out = {}
for name, x in (("c64_gauss", torch.randn((B, N, 64), device="cuda")), ("c3_ball", torch.from_numpy(make_clouds("ball", B, N, seed=5)).cuda()),
                ("c64_relu", torch.relu(torch.randn((B, N, 64), device="cuda"))),
                # a feature cloud far from the origin (|x|^2 ~ 50 x the neighbour distances), like DGCNN's post-ReLU EdgeConv outputs
                ("c64_offset", torch.relu(torch.randn((B, N, 64), device="cuda") * 0.3 + 2.0))):
    c = x.shape[2]
    need = lib.psa_knn_graph_workspace_bytes(B, N, c, K)
    ws = torch.zeros(need // 4 + 1, dtype=torch.float32, device="cuda")
    idx = torch.empty((B, N, K), dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())
    run = lambda: lib.psa_knn_graph_ws(B, N, c, K, vp(x), vp(idx), vp(ws), C.c_size_t(need), st)
    for _ in range(3):
        assert run() == 0
    torch.cuda.synchronize()
    npad = (N + 127) // 128 * 128
    off = (B * (npad // 128) * 49152 + ((B * npad * 4 + 255) & ~255) + ((B * N * 4 + 255) & ~255)) // 4
    flagged = int(ws[off:off + 1].view(torch.int32).item())
    # PSA_KNN_ERRSTAT build only (python tools/build_variant.py errstat -DPSA_KNN_ERRSTAT; PSA_LIB_PATH=...): the largest observed
    # |fine - canonical| in units of the bound E2 and the number of canonically evaluated (ambiguous) entries; zeros otherwise
    err_over_e2 = float(ws[off + 1:off + 2].view(torch.float32).item())
    ambiguous = int(ws[off + 2:off + 3].view(torch.int32).item())
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    ops._KNN_FP32_ONLY = True
    for _ in range(2):
        ref = ops.knn_graph(x, K)
    t2 = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ref = ops.knn_graph(x, K); e1.record(); torch.cuda.synchronize()
        t2.append(e0.elapsed_time(e1) * 1e3)
    t2.sort()
    ops._KNN_FP32_ONLY = False
    out[name] = {"max_err_over_E2": err_over_e2, "ambiguous_entries_per_row": ambiguous / (B * N), "tc_us": ts[len(ts) // 2], "fp32_us": t2[len(t2) // 2], "rows_to_exhaustive_kernel": flagged, "rows": B * N,
                 "equal": bool(torch.equal(idx, ref))}
print(json.dumps(out))
