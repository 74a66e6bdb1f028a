"""Variant F1 at the SA1 shape of the bench (B=32, N=2048, m=512, K=32, C1=64): isolated (L2 read-flushed) and steady-state
(back-to-back) time of psa_sa_conv1_prebn through the C ABI with preallocated buffers (no allocator, one ctypes call per launch).
PSA_F1_VARIANT=1 selects the round-1 kernel.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from scanobjectnn_b200 import _lib, ops, pointnet2_cls_ssg
from scanobjectnn_b200.synthetic import make_clouds

B, N, M, K, C1 = 32, 2048, 512, 32, 64
if len(sys.argv) > 1 and sys.argv[1] == "sa2":
    B, N, M, K, C1 = 32, 512, 128, 64, 128
lib = _lib.load()
p = pointnet2_cls_ssg.init_params(seed=1, randomize_bn=True)
x = torch.from_numpy(make_clouds("ball", B, N, seed=1001)).cuda()
_, l1 = ops.farthest_point_sample_and_gather(M, x)
feats = torch.randn((B, N, 128), device="cuda") if C1 == 128 else None
c = 128 if feats is not None else 0
w1 = (torch.randn((3 + c, C1), device="cuda") * 0.1).contiguous()
bias = torch.randn(C1, device="cuda") * 0.1
pre = torch.empty((B, M, K, C1), device="cuda")
idx = torch.empty((B, M, K), dtype=torch.int32, device="cuda")
cnt = torch.empty((B, M), dtype=torch.int32, device="cuda")
stats = torch.empty((2, C1), device="cuda")
need = lib.psa_sa_conv1_prebn_workspace_bytes(B, N, M, c, C1, 1)
ws = torch.empty(need // 4 + 1, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def launch():
    rc = lib.psa_sa_conv1_prebn(B, N, M, c, C.c_float(0.2 if C1 == 64 else 0.4), K, vp(x), vp(l1), vp(feats), vp(w1), vp(bias), C1, vp(pre), vp(idx),
                                vp(cnt), vp(stats), vp(ws), C.c_size_t(need), st)
    assert rc == 0, ops._lib.last_error() if hasattr(ops._lib, "last_error") else rc


for _ in range(3):
    launch()
torch.cuda.synchronize()
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
sink = torch.zeros((), device="cuda")
iso = []
for _ in range(15):
    sink.copy_(flush.sum())
    torch.cuda._sleep(400_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(); e1.record()
    torch.cuda.synchronize()
    iso.append(e0.elapsed_time(e1) * 1e3)
iso.sort()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(2_000_000)
e0.record()
for _ in range(50):
    launch()
e1.record()
torch.cuda.synchronize()
steady = e0.elapsed_time(e1) * 1e3 / 50
tl = None
if os.environ.get("PSA_F1_TLOG"):
    # instrumented build (tools/build_variant.py f1timing -DPSA_F1_TIMING): per-CTA globaltimer stamps behind the CTA partials
    torch.cuda.synchronize()
    launch()
    torch.cuda.synchronize()
    off = (256 + 3 * 148 * 2 * C1 * 4) // 4
    allraw = ws[off:off + 444 * 24 * 2].view(torch.int64).cpu().numpy()
    raw = allraw[:444 * 8].reshape(-1, 8)
    keep = raw[:, 0] > 0
    ph = allraw[444 * 8:444 * 24].reshape(-1, 16)[keep]   # cycles per stage (barrier-ordered clock64 reads), summed over the CTA's batches
    raw = raw[keep]
    if os.environ.get("PSA_F1_TLOG_DUMP"):
        import numpy as np
        np.save(os.environ["PSA_F1_TLOG_DUMP"], raw)
    t0 = raw[:, 0].min()
    med = lambda c: float(sorted(raw[:, c] - t0)[len(raw) // 2]) / 1e3   # noqa: E731
    mx = lambda c: float((raw[:, c] - t0).max()) / 1e3                   # noqa: E731
    tl = {"ctas": int(len(raw)), "us_from_first_cta_start": {"cta_start_max": mx(0), "cloud_loaded_median": med(1), "cloud_loaded_max": mx(1),
          "first_search_done_median": med(4), "first_rows_ready_median": med(2), "first_rows_ready_max": mx(2),
          "consumers_done_median": med(5), "consumers_done_max": mx(5), "cta_end_max": mx(6)},
          "kcycles_per_cta_median": {k: float(sorted(ph[:, c])[len(ph) // 2]) / 1e3 for c, k in
                                     [(1, "search_wait_space"), (2, "search_work"), (3, "extract_wait_bitmaps"), (4, "extract_wait_space"), (5, "extract_work"),
                                      (6, "conv_wait_rows"), (8, "conv_work")]}}
nbytes = B * (12 * N + 12 * M) + 4 * B * M * K * C1 + 4 * B * M * K + 4 * B * M + (4 * B * N * C1 if c else 0)
print(json.dumps({"variant": os.environ.get("PSA_F1_VARIANT", "0"), "shape": [B, N, M, K, C1, c], "alg_bytes": nbytes,
                  "isolated_us_median": iso[len(iso) // 2], "isolated_us_min": iso[0], "isolated_gbs": nbytes / iso[len(iso) // 2] / 1e3,
                  "steady_us": steady, "steady_gbs": nbytes / steady / 1e3, "timeline": tl}))
