"""The conv/BN/ReLU/max half of the hot path lives in TensorFlow in the reference (not installable here: "parity unpinned").
Second opinion for the numpy restatement (oracle/mlp_oracle.py): the same layers evaluated by an independent library
implementation -- torch.nn.functional conv2d (1x1 kernel, NHWC data moved to NCHW), batch_norm in inference mode with
TF's epsilon 1e-3, relu, amax over the nsample axis -- in float64 on CPU, on the tensors sample_and_group produces."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import mlp_oracle as mo
from scanobjectnn_b200.pointnet_util import add_sa_module_params
from scanobjectnn_b200.synthetic import make_clouds
from scanobjectnn_b200.tf_util import VariableStore


def _torch_layer(x_nhwc, p, scope, relu=True):
    w = p[f"{scope}/weights"].double()                      # TF kernel (1,1,Cin,Cout) stored as (Cin,Cout)
    w = w.reshape(-1, w.shape[-1]).t()[:, :, None, None]    # -> torch (Cout,Cin,1,1)
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w, bias=p[f"{scope}/biases"].double())
    if f"{scope}/bn/gamma" in p:
        y = F.batch_norm(y, p[f"{scope}/bn/moving_mean"].double(), p[f"{scope}/bn/moving_variance"].double(),
                         p[f"{scope}/bn/gamma"].double(), p[f"{scope}/bn/beta"].double(), training=False, eps=1e-3)
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1)


def test_sa_module_restatement_agrees_with_torch_functional():
    p = VariableStore(device="cpu", seed=3)
    mlp = [32, 48, 96]
    add_sa_module_params(p, "sa", 3 + 5, mlp, randomize_bn=True)
    xyz = make_clouds("shell", 2, 256, seed=77)
    pts = np.random.default_rng(1).standard_normal((2, 256, 5)).astype(np.float32)
    new_xyz, want, idx = mo.sa_module(xyz, pts, 32, 0.3, 16, mlp, False, "sa", p)
    # the grouped tensor exactly as pointnet_util.sample_and_group builds it: [xyz - centre, features], xyz first
    b = np.arange(2)[:, None, None]
    gx = xyz[b, idx] - new_xyz[:, :, None, :]
    g = torch.from_numpy(np.concatenate([gx, pts[b, idx]], axis=-1)).double()          # (B,m,K,3+C)
    for i in range(len(mlp)):
        g = _torch_layer(g, p, f"sa/conv{i}")
    got = g.amax(dim=2).numpy()
    assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())


def test_dense_chain_restatement_agrees_with_torch_functional():
    p = VariableStore(device="cpu", seed=4)
    p.add_conv2d("c0", 7, 64, randomize_bn=True)
    p.add_conv2d("c1", 64, 33, bn=False)
    x = np.random.default_rng(2).standard_normal((3, 50, 1, 7)).astype(np.float32)
    want = mo.mlp_chain(x, p, ["c0", "c1"], [True, False])
    g = _torch_layer(torch.from_numpy(x).double(), p, "c0")
    g = _torch_layer(g, p, "c1", relu=False)
    assert np.abs(g.numpy() - want).max() < 1e-12 * max(1.0, np.abs(want).max())


def test_host_speed_fp32_forward_agrees_with_fp64_restatement():
    """bench.py's CPU legs time mlp_oracle.pointnet2_cls_ssg_fast (one fp32 GEMM per layer, folded BN): same logits as the
    float64 restatement of pointnet2_cls_ssg.py:23-47 to fp32 rounding."""
    from scanobjectnn_b200 import pointnet2_cls_ssg

    p = pointnet2_cls_ssg.init_params(seed=5, device="cpu", randomize_bn=True)
    xyz = make_clouds("shell", 2, 1024, seed=77)
    want, _ = mo.pointnet2_cls_ssg(xyz, p)
    got = mo.pointnet2_cls_ssg_fast(xyz, p, threads=2)
    assert got.shape == want.shape == (2, 15)
    assert np.abs(got - want).max() < 1e-5 * max(1.0, np.abs(want).max())
