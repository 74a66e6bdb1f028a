"""Event-timed psa_shared_mlp layers at the SA3 shapes (rows = 4096): one launch per layer, L2-warm and L2-cold."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scanobjectnn_b200 import ops

dev = "cuda"
flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
sink = torch.zeros((), device=dev)
for rows, K, N, pk in [(4096, 256, 256, 1), (4096, 256, 512, 1), (4096, 512, 1024, 128), (16384, 128, 128, 1), (65536, 64, 128, 1), (65536, 128, 1024, 2048)]:
    x = torch.randn((rows, K), device=dev)
    mlp = ops.MlpParams([(torch.randn((K, N), device=dev) * 0.05, torch.ones(N, device=dev), torch.zeros(N, device=dev), True)])
    fn = lambda: ops.shared_mlp(x, mlp, pool_k=pk)
    for _ in range(3): fn()
    res = {}
    for cold in (False, True):
        ts = []
        for _ in range(7):
            if cold: sink.copy_(flush.sum())
            torch.cuda._sleep(300_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort(); res["cold" if cold else "warm"] = ts[len(ts) // 2]
    fl = 2.0 * rows * K * N
    print(f"rows={rows} {K}->{N} pool={pk}: warm {res['warm']:.1f} us ({fl / res['warm'] / 1e6:.1f} TFLOP/s fp32-eq)  cold {res['cold']:.1f} us")

# with the -DPSA_TC_TIMING build: cycles of row-warp 0 / thread 0 per CTA in each phase of tc_dense2_kernel
import ctypes as C
from scanobjectnn_b200 import _lib
lib = _lib.load()
if hasattr(lib, "psa_debug_tc_timing"):
    fn_t = lib.psa_debug_tc_timing
    fn_t.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    names = ["setup+first x load", "wait stage free", "split+store+next load", "final mma wait", "epilogue", "final_sync", "-"]
    for rows, K, N, pk in [(4096, 256, 256, 1), (4096, 512, 1024, 128)]:
        x = torch.randn((rows, K), device=dev)
        mlp = ops.MlpParams([(torch.randn((K, N), device=dev) * 0.05, torch.ones(N, device=dev), torch.zeros(N, device=dev), True)])
        ops.shared_mlp(x, mlp, pool_k=pk); ops.shared_mlp(x, mlp, pool_k=pk)
        fn_t(None, 1)
        ops.shared_mlp(x, mlp, pool_k=pk)
        out = (C.c_ulonglong * 8)(); fn_t(out, 1)
        t = list(out); n = max(1, t[7])
        print(f"rows={rows} {K}->{N}: CTAs {t[7]} cycles/CTA", {k: round(v / n) for k, v in zip(names, t[:7])})
