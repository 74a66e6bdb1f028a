"""CPU restatement of the TF-library half of the hot path: 1x1 conv + bias + batch norm (inference) + ReLU +
max-pool, and the model assemblies built from it.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

PARITY UNPINNED: this arithmetic lives in TensorFlow 1.x (cuDNN conv, contrib batch_norm), which is not part of
/root/reference and not installable here; the reference holds no golden vectors for it.  What is restated is the
published semantics at the reference's call sites:
  conv2d 1x1 + bias_add          pointnet2/utils/tf_util.py:155-176  (NHWC matmul over the channel axis)
  batch_norm, is_training=False  pointnet2/utils/tf_util.py:512-531  -> (x-mean)*gamma*rsqrt(var+1e-3)+beta
  relu                           pointnet2/utils/tf_util.py:183-184
  reduce_max over nsample        pointnet2/utils/pointnet_util.py:127
The fp64 evaluation is the truth the CUDA path is held to (1e-5); the fp32 one is the "plain fp32" comparison.
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc

BN_EPS = 1e-3


def _np(t, dtype):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=dtype)


def conv_bn_relu(x, params, scope, relu=True, dtype=np.float64):
    """x (..., Cin) -> (..., Cout) for the layer stored under ``scope`` (TF variable names)."""
    w = _np(params[f"{scope}/weights"], dtype)
    w = w.reshape(-1, w.shape[-1])
    y = x.astype(dtype) @ w + _np(params[f"{scope}/biases"], dtype)
    if f"{scope}/bn/gamma" in params:
        inv = _np(params[f"{scope}/bn/gamma"], dtype) / np.sqrt(_np(params[f"{scope}/bn/moving_variance"], dtype) + dtype(BN_EPS))
        y = (y - _np(params[f"{scope}/bn/moving_mean"], dtype)) * inv + _np(params[f"{scope}/bn/beta"], dtype)
    if relu:
        y = np.maximum(y, 0)
    return y


def mlp_chain(x, params, scopes, relus=None, dtype=np.float64):
    relus = relus or [True] * len(scopes)
    for s, r in zip(scopes, relus):
        x = conv_bn_relu(x, params, s, r, dtype)
    return x


def sample_and_group(npoint, radius, nsample, xyz, points):
    """pointnet_util.sample_and_group (pointnet_util.py:22-56) on the C oracle ops."""
    idx_fps = orc.fps(xyz, npoint)
    new_xyz = orc.gather_point(xyz, idx_fps)
    idx, cnt = orc.query_ball_point(radius, nsample, xyz, new_xyz, contract=True)
    grouped_xyz = orc.group_point(xyz, idx) - new_xyz[:, :, None, :]
    if points is not None:
        new_points = np.concatenate([grouped_xyz, orc.group_point(points, idx)], axis=-1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, idx_fps


def sa_module(xyz, points, npoint, radius, nsample, mlp, group_all, scope, params, dtype=np.float64, pooling="max", mlp2=None):
    """pointnet_sa_module (pointnet_util.py:87-154), inference; pooling modes :126-146, post-MLP :148-157."""
    scopes = [f"{scope}/conv{i}" for i in range(len(mlp))]
    if group_all:
        new_xyz = np.zeros((xyz.shape[0], 1, 3), np.float32)
        new_points = (np.concatenate([xyz, points], axis=2) if points is not None else xyz)[:, None]
        grouped_xyz = xyz[:, None]
        idx = None
    else:
        new_xyz, new_points, idx, _ = sample_and_group(npoint, radius, nsample, xyz, points)
        grouped_xyz = new_points[..., :3]
    y = mlp_chain(new_points, params, scopes, dtype=dtype)
    if pooling == "max":
        out = y.max(axis=2)
    elif pooling == "avg":
        out = y.mean(axis=2)
    elif pooling == "weighted_avg":
        w = np.exp(-np.linalg.norm(grouped_xyz.astype(dtype), axis=-1, keepdims=True) * 5)
        out = (y * (w / w.sum(axis=2, keepdims=True))).sum(axis=2)
    elif pooling == "max_and_avg":
        out = np.concatenate([y.mean(axis=2), y.max(axis=2)], axis=-1)
    else:
        raise ValueError(pooling)
    if mlp2 is not None:
        out = mlp_chain(out, params, [f"{scope}/conv_post_{i}" for i in range(len(mlp2))], dtype=dtype)
    return new_xyz, out.astype(dtype), idx


def sa_module_msg(xyz, points, npoint, radius_list, nsample_list, mlp_list, scope, params, dtype=np.float64):
    """pointnet_sa_module_msg (pointnet_util.py:156-196): per scale ball query, rows [features, xyz] (xyz LAST, :184),
    MLP conv{i}_{j}, max over nsample; concat over scales."""
    new_xyz = orc.gather_point(xyz, orc.fps(xyz, npoint))
    outs = []
    for i, (radius, nsample, mlp) in enumerate(zip(radius_list, nsample_list, mlp_list)):
        idx, _ = orc.query_ball_point(radius, nsample, xyz, new_xyz, contract=True)
        g = orc.group_point(xyz, idx) - new_xyz[:, :, None, :]
        rows = g if points is None else np.concatenate([orc.group_point(points, idx), g], axis=-1)
        y = mlp_chain(rows, params, [f"{scope}/conv{i}_{j}" for j in range(len(mlp))], dtype=dtype)
        outs.append(y.max(axis=2))
    return new_xyz, np.concatenate(outs, axis=-1).astype(dtype)


def fp_module(xyz1, xyz2, points1, points2, mlp, scope, params, dtype=np.float64):
    """pointnet_fp_module (pointnet_util.py:199-229), inference."""
    dist, idx = orc.three_nn(xyz1, xyz2)
    w = orc.three_weights(dist)
    interp = orc.three_interpolate(np.asarray(points2, np.float32), idx, w)
    x = np.concatenate([interp, points1], axis=2) if points1 is not None else interp
    return mlp_chain(x, params, [f"{scope}/conv_{i}" for i in range(len(mlp))], dtype=dtype)


def pointnet2_cls_ssg(point_cloud, params, dtype=np.float64):
    """pointnet2/models/pointnet2_cls_ssg.py:23-47, is_training=False -> logits, end_points."""
    xyz = np.asarray(point_cloud, np.float32)
    l1_xyz, l1_points, l1_idx = sa_module(xyz, None, 512, 0.2, 32, [64, 64, 128], False, "layer1", params, dtype)
    l2_xyz, l2_points, l2_idx = sa_module(l1_xyz, l1_points.astype(np.float32), 128, 0.4, 64, [128, 128, 256], False,
                                          "layer2", params, dtype)
    _, l3_points, _ = sa_module(l2_xyz, l2_points.astype(np.float32), None, None, None, [256, 512, 1024], True,
                                "layer3", params, dtype)
    net = l3_points.reshape(xyz.shape[0], -1)
    net = mlp_chain(net, params, ["fc1", "fc2", "fc3"], [True, True, False], dtype)
    return net, dict(l1_xyz=l1_xyz, l1_points=l1_points, l1_idx=l1_idx, l2_xyz=l2_xyz, l2_points=l2_points,
                     l2_idx=l2_idx, l3_points=l3_points)


def edgeconv(x, nn_idx, params, scopes, dtype=np.float64):
    """get_edge_feature + conv chain + max over k (dgcnn/utils/tf_util.py:674-706, dgcnn.py:41-47)."""
    x = np.asarray(x, np.float32)
    b, n, c = x.shape
    nb = x[np.arange(b)[:, None, None], np.asarray(nn_idx, np.int64)]
    ctr = np.broadcast_to(x[:, :, None, :], nb.shape)
    edge = np.concatenate([ctr, nb - ctr], axis=-1)
    return mlp_chain(edge, params, scopes, dtype=dtype).max(axis=2)


def pointnet2_cls_bga(point_cloud, params, dtype=np.float64):
    """pointnet2/models/pointnet2_cls_bga.py:21-75, is_training=False -> class_pred, seg_pred."""
    xyz = np.asarray(point_cloud, np.float32)[:, :, :3]
    b = xyz.shape[0]
    l1_xyz, l1_points, _ = sa_module(xyz, None, 512, 0.2, 64, [64, 64, 128], False, "layer1", params, dtype)
    l2_xyz, l2_points, _ = sa_module(l1_xyz, l1_points.astype(np.float32), 128, 0.4, 64, [128, 128, 256], False, "layer2", params, dtype)
    l3_xyz, l3_points, _ = sa_module(l2_xyz, l2_points.astype(np.float32), None, None, None, [256, 512, 1024], True, "layer3", params, dtype)
    net = mlp_chain(l3_points.reshape(b, -1), params, ["fc1", "fc2"], dtype=dtype)
    class_vector = net[:, None, :]
    class_pred = mlp_chain(net, params, ["fc3"], [False], dtype)
    l2p = fp_module(l2_xyz, l3_xyz, l2_points.astype(np.float32), class_vector.astype(np.float32), [256, 256], "fa_layer1", params, dtype)
    l1p = fp_module(l1_xyz, l2_xyz, l1_points.astype(np.float32), l2p.astype(np.float32), [256, 128], "fa_layer2", params, dtype)
    l0p = fp_module(xyz, l1_xyz, None, l1p.astype(np.float32), [128, 128, 128], "fa_layer3", params, dtype)
    feats = mlp_chain(l0p, params, ["seg_fc1"], dtype=dtype)
    seg = mlp_chain(feats, params, ["seg_fc2"], [False], dtype)
    return class_pred, seg


def dgcnn_stage(x, k, params, scopes, dtype=np.float64):
    """knn graph (canonical fp32 order, C oracle) + EdgeConv + max over k, on the given fp32 features."""
    x = np.asarray(x, np.float32)
    idx = orc.dgcnn_knn(x, k)
    return idx, edgeconv(x, idx, params, scopes, dtype)


def _tnet(x, params, scope, K, dtype):
    g = mlp_chain(x, params, [f"{scope}/tconv1", f"{scope}/tconv2", f"{scope}/tconv3"], dtype=dtype).max(axis=1)
    g = mlp_chain(g, params, [f"{scope}/tfc1", f"{scope}/tfc2"], dtype=dtype)
    name = "transform_XYZ" if K == 3 else "transform_feat"
    w = _np(params[f"{scope}/{name}/weights"], dtype)
    b = _np(params[f"{scope}/{name}/biases"], dtype) + np.eye(K, dtype=dtype).flatten()
    return (g @ w + b).reshape(-1, K, K)


def pointnet_cls(point_cloud, params, dtype=np.float64):
    """pointnet/models/pointnet_cls.py:21-75, is_training=False -> logits, feature transform."""
    x = np.asarray(point_cloud, np.float32).astype(dtype)
    t1 = _tnet(x, params, "transform_net1", 3, dtype)
    x = x @ t1
    net = mlp_chain(x, params, ["conv1", "conv2"], dtype=dtype)
    t2 = _tnet(net, params, "transform_net2", 64, dtype)
    net = net @ t2
    net = mlp_chain(net, params, ["conv3", "conv4", "conv5"], dtype=dtype).max(axis=1)
    return mlp_chain(net, params, ["fc1", "fc2", "fc3"], [True, True, False], dtype), t2


# ---------------------------------------------------------------------------------------------------------------------
# Host-speed fp32 evaluation of the same forward, for bench.py's CPU legs only (cpu_baseline, --impl reference).
# Same semantics as pointnet2_cls_ssg above; organised the way a CPU build of the reference would run it (SURVEY 8d:
# "grouped MLP via PyTorch-CPU fp32 with torch.set_num_threads(nproc)"): every 1x1 conv is ONE (rows x Cin) x (Cin x Cout)
# GEMM on all host cores (oneDNN / MKL through torch), the inference batch norm is folded into the weights and the bias
# (folded in float64, then rounded), ReLU in place, max over nsample as one reduction -- instead of numpy's batched 4-D matmul
# (one small GEMM per neighbourhood) and separate elementwise passes.  tests/test_mlp_oracle_crosscheck.py holds it to the
# fp64 evaluation.
# ---------------------------------------------------------------------------------------------------------------------
def _folded(params, scope):
    w = _np(params[f"{scope}/weights"], np.float64)
    w = w.reshape(-1, w.shape[-1])
    b = _np(params[f"{scope}/biases"], np.float64)
    if f"{scope}/bn/gamma" in params:
        inv = _np(params[f"{scope}/bn/gamma"], np.float64) / np.sqrt(_np(params[f"{scope}/bn/moving_variance"], np.float64) + BN_EPS)
        w = w * inv
        b = (b - _np(params[f"{scope}/bn/moving_mean"], np.float64)) * inv + _np(params[f"{scope}/bn/beta"], np.float64)
    return w.astype(np.float32), b.astype(np.float32)


_FAST_WS = {}


def _flat(i, numel):
    """Reusable flat fp32 buffer number i (grown on demand): a fresh 100-300 MB output per layer costs more in page faults than
    the GEMM that fills it (measured here: 45-400 ms fresh vs 18 ms into a resident buffer for 524288 x 64 x 128)."""
    import torch

    t = _FAST_WS.get(i)
    if t is None or t.numel() < numel:
        t = torch.empty(int(numel), dtype=torch.float32)
        t.zero_()                                  # touch the pages once
        _FAST_WS[i] = t
    return t[:numel]


def _layer_fast(out_flat, x, params, scope, relu=True, x2=None, split=None):
    """relu(x . W' + b') with the folded weights of ``scope`` into a view of ``out_flat``; with x2/split the input is the
    virtual concatenation [x2, x] (W' rows [0, split) act on x2) -- the concat tensor of pointnet_util.py:50 is never built."""
    import torch

    w, b = _folded(params, scope)
    wt, bt = torch.from_numpy(w), torch.from_numpy(b)
    out = out_flat[: x.shape[0] * w.shape[1]].view(x.shape[0], w.shape[1])
    if x2 is None:
        torch.addmm(bt, x, wt, out=out)
    else:
        torch.addmm(bt, x, wt[split:], out=out)
        out.addmm_(x2, wt[:split])
    if relu:
        out.relu_()
    return out


def _mlp_rows_fast(rows, params, scopes, relus=None):
    """rows: torch fp32 (R, Cin) -> (R, Cout); small layers (group_all level, FC head): fresh outputs."""
    import torch

    relus = relus or [True] * len(scopes)
    for s_, r in zip(scopes, relus):
        w, b = _folded(params, s_)
        rows = torch.addmm(torch.from_numpy(b), rows, torch.from_numpy(w))
        if r:
            rows.relu_()
    return rows


def pointnet2_cls_ssg_fast(point_cloud, params, threads=None):
    """pointnet2_cls_ssg (pointnet2/models/pointnet2_cls_ssg.py:23-47, is_training=False) in fp32 at host speed -> logits.
    Index ops: the C restatement (OpenMP over the batch); grouped MLPs: one GEMM per layer into reused buffers."""
    import os

    import torch

    torch.set_num_threads(threads or os.cpu_count() or 1)
    xyz = np.ascontiguousarray(point_cloud, np.float32)
    bsz = xyz.shape[0]
    lib = orc.lib()
    with torch.no_grad():
        cur_xyz, cur_pts = xyz, None
        for scope, npoint, radius, nsample in (("layer1", 512, 0.2, 32), ("layer2", 128, 0.4, 64)):
            new_xyz = orc.gather_point(cur_xyz, orc.fps(cur_xyz, npoint))
            idx, _ = orc.query_ball_point(radius, nsample, cur_xyz, new_xyz, contract=True)
            rows = bsz * npoint * nsample
            gx = torch.from_numpy((orc.group_point(cur_xyz, idx) - new_xyz[:, :, None, :]).reshape(rows, 3))
            widths = [_folded(params, f"{scope}/conv{i}")[0].shape[1] for i in range(3)]
            if cur_pts is None:
                y = _layer_fast(_flat(0, rows * widths[0]), gx, params, f"{scope}/conv0")
                nxt = 1
            else:
                c = cur_pts.shape[-1]
                gf = _flat(0, rows * c).view(rows, c)                      # grouped features straight into the reused buffer
                lib.orc_group_point(bsz, cur_pts.shape[1], c, npoint, nsample, orc._p(np.ascontiguousarray(cur_pts, np.float32)), orc._p(idx),
                                    orc._p(gf.numpy()))
                y = _layer_fast(_flat(1, rows * widths[0]), gf, params, f"{scope}/conv0", x2=gx, split=3)
                nxt = 0
            y = _layer_fast(_flat(nxt, rows * widths[1]), y, params, f"{scope}/conv1")
            y = _layer_fast(_flat(2, rows * widths[2]), y, params, f"{scope}/conv2")
            cur_pts = y.view(bsz * npoint, nsample, widths[2]).amax(dim=1).view(bsz, npoint, widths[2]).numpy()
            cur_xyz = new_xyz
        rows = torch.from_numpy(np.concatenate([cur_xyz, cur_pts], axis=2)).reshape(-1, 3 + cur_pts.shape[-1])
        y = _mlp_rows_fast(rows, params, [f"layer3/conv{i}" for i in range(3)])
        net = y.reshape(bsz, -1, y.shape[-1]).amax(dim=1)
        net = _mlp_rows_fast(net, params, ["fc1", "fc2", "fc3"], [True, True, False])
    return net.numpy()
