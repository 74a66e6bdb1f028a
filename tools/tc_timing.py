"""Where does a tile of tc_sa_kernel spend its time?  Needs the debug build:
    PSA_EXTRA_NVCC_FLAGS=-DPSA_TC_TIMING python -m scanobjectnn_b200.build --force
Prints, per SA level, thread 0's average cycles per tile in each phase (layer-1 gather+store, MMA issue, MMA wait,
epilogue, for the inner and the last tensor layer)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from scanobjectnn_b200 import _lib, ops, pointnet2_cls_ssg
from scanobjectnn_b200.synthetic import make_clouds

lib = _lib.load()
fn = lib.psa_debug_tc_timing
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
params = pointnet2_cls_ssg.init_params(seed=1, randomize_bn=True)
x = torch.from_numpy(make_clouds("ball", 32, 2048, seed=1001)).cuda()
_, l1 = ops.farthest_point_sample_and_gather(512, x)
_, l2 = ops.farthest_point_sample_and_gather(128, l1)
SC = lambda sc: [f"{sc}/conv{i}" for i in range(3)]
names = ["layer1", "mma_issue", "mma_wait", "epilogue", "last_issue", "last_wait", "last_epilogue", "tiles"]
cases = {
    "sa1": lambda: ops.sa_module_infer(x, l1, None, 0.2, 32, params.mlp(SC("layer1"))),
}
f1 = cases["sa1"]()
cases["sa2"] = lambda: ops.sa_module_infer(l1, l2, f1, 0.4, 64, params.mlp(SC("layer2")))
for name, call in cases.items():
    call(); call()
    fn(None, 1)
    call()
    out = (C.c_ulonglong * 8)()
    fn(out, 1)
    t = list(out)
    tiles = max(1, t[7])
    tot = sum(t[:7])
    print(name, "tiles", t[7], "cycles/tile", round(tot / tiles), {n: round(v / tiles) for n, v in zip(names[:7], t[:7])})
