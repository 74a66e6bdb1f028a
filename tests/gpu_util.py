"""Helpers for the -m gpu parity tests: torch<->numpy moves and ctypes calls into the compiled REFERENCE CUDA
kernels (oracle/_ref/libref_tfops.so) with torch device pointers."""
import ctypes as C

import numpy as np
import torch

from oracle import oracle as orc


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def npy(t):
    return t.detach().cpu().numpy()


def _p(t):
    return C.c_void_p(t.data_ptr())


def ref_fps(xyz_t, m):
    b, n, _ = xyz_t.shape
    temp = torch.empty((32, n), dtype=torch.float32, device="cuda")
    out = torch.zeros((b, m), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rc = orc.refgpu().ref_fps(b, n, m, _p(xyz_t), _p(temp), _p(out), 1)
    assert rc == 0, rc
    return out


def ref_query_ball_point(radius, nsample, xyz1_t, xyz2_t, fill=0):
    b, n, _ = xyz1_t.shape
    m = xyz2_t.shape[1]
    idx = torch.full((b, m, nsample), fill, dtype=torch.int32, device="cuda")
    cnt = torch.zeros((b, m), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rc = orc.refgpu().ref_query_ball_point(b, n, m, C.c_float(radius), nsample, _p(xyz1_t), _p(xyz2_t), _p(idx), _p(cnt), 1)
    assert rc == 0, rc
    return idx, cnt


def ref_group_point(points_t, idx_t):
    b, n, c = points_t.shape
    _, m, k = idx_t.shape
    out = torch.empty((b, m, k, c), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rc = orc.refgpu().ref_group_point(b, n, c, m, k, _p(points_t), _p(idx_t), _p(out), 1)
    assert rc == 0, rc
    return out


def ref_selection_sort(k, dist_t):
    b, m, n = dist_t.shape
    outi = torch.empty((b, m, n), dtype=torch.int32, device="cuda")
    out = torch.empty((b, m, n), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rc = orc.refgpu().ref_selection_sort(b, n, m, k, _p(dist_t), _p(outi), _p(out), 1)
    assert rc == 0, rc
    return outi, out


def ref_gather_point(inp_t, idx_t):
    b, n, _ = inp_t.shape
    m = idx_t.shape[1]
    out = torch.empty((b, m, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rc = orc.refgpu().ref_gather_point(b, n, m, _p(inp_t), _p(idx_t), _p(out), 1)
    assert rc == 0, rc
    return out
