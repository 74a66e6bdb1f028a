"""pointnet/models/pointnet_cls.py (vanilla PointNet, BASELINE.json configs[0]: the reference's plumbing case) on the same
dense kernels: every layer is a per-point shared MLP (`psa_shared_mlp`), the symmetric function is its fused max-pool.
get_model(point_cloud, is_training, bn_decay, num_class) -> (logits (B,num_class), end_points).  Inference mode."""
from __future__ import annotations

import torch

from . import ops
from .tf_util import VariableStore, _require_inference

NUM_CLASSES = 15


def _add_tnet(p: VariableStore, scope: str, cin: int, K: int, rb: bool):
    p.add_conv2d(f"{scope}/tconv1", cin, 64, randomize_bn=rb)        # conv [1,3] over (B,N,3,1) == 3 -> 64 per point
    p.add_conv2d(f"{scope}/tconv2", 64, 128, randomize_bn=rb)
    p.add_conv2d(f"{scope}/tconv3", 128, 1024, randomize_bn=rb)
    p.add_fc(f"{scope}/tfc1", 1024, 512, randomize_bn=rb)
    p.add_fc(f"{scope}/tfc2", 512, 256, randomize_bn=rb)
    name = "transform_XYZ" if K == 3 else "transform_feat"            # transform_nets.py:38,80
    p[f"{scope}/{name}/weights"] = torch.zeros((256, K * K), device=p.device)
    p[f"{scope}/{name}/biases"] = torch.zeros(K * K, device=p.device)


def init_params(num_class=NUM_CLASSES, seed=0, device="cuda", randomize_bn=False) -> VariableStore:
    p = VariableStore(device=device, seed=seed)
    _add_tnet(p, "transform_net1", 3, 3, randomize_bn)
    p.add_conv2d("conv1", 3, 64, randomize_bn=randomize_bn)
    p.add_conv2d("conv2", 64, 64, randomize_bn=randomize_bn)
    _add_tnet(p, "transform_net2", 64, 64, randomize_bn)
    p.add_conv2d("conv3", 64, 64, randomize_bn=randomize_bn)
    p.add_conv2d("conv4", 64, 128, randomize_bn=randomize_bn)
    p.add_conv2d("conv5", 128, 1024, randomize_bn=randomize_bn)
    p.add_fc("fc1", 1024, 512, randomize_bn=randomize_bn)
    p.add_fc("fc2", 512, 256, randomize_bn=randomize_bn)
    p.add_fc("fc3", 256, num_class, bn=False)
    return p


def transform_net(x, params: VariableStore, scope: str, K: int):
    """input_transform_net / feature_transform_net (pointnet/models/transform_nets.py:10-97): (B,N,C) -> (B,K,K)."""
    b, n, c = x.shape
    g = ops.shared_mlp(x.reshape(b * n, c), params.mlp([f"{scope}/tconv1", f"{scope}/tconv2", f"{scope}/tconv3"]), pool_k=n)
    g = ops.shared_mlp(g, params.mlp([f"{scope}/tfc1", f"{scope}/tfc2"]))
    name = "transform_XYZ" if K == 3 else "transform_feat"
    w = params[f"{scope}/{name}/weights"]
    bias = params[f"{scope}/{name}/biases"] + torch.eye(K, device=w.device).flatten()
    return (g @ w + bias).reshape(b, K, K)


def get_model(point_cloud, is_training, bn_decay=None, num_class=NUM_CLASSES, *, params: VariableStore):
    _require_inference(is_training)
    b, n, _ = point_cloud.shape
    end_points = {}
    t1 = transform_net(point_cloud, params, "transform_net1", 3)
    x = torch.bmm(point_cloud, t1).contiguous()                                       # tf.matmul(point_cloud, transform)
    net = ops.shared_mlp(x.reshape(b * n, 3), params.mlp(["conv1", "conv2"])).reshape(b, n, 64)
    t2 = transform_net(net, params, "transform_net2", 64)
    end_points["transform"] = t2
    net = torch.bmm(net, t2).contiguous()
    net = ops.shared_mlp(net.reshape(b * n, 64), params.mlp(["conv3", "conv4", "conv5"]), pool_k=n)   # max over the N points
    end_points["global"] = net
    net = ops.shared_mlp(net, params.mlp(["fc1", "fc2", "fc3"], [True, True, False]))
    return net, end_points


def get_loss(pred, label, end_points, reg_weight=0.001):
    """classification CE + reg_weight * || I - T T^t ||_F^2 / 2 on the feature transform (pointnet_cls.py:78-95)."""
    ce = torch.nn.functional.cross_entropy(pred, label.long())
    t = end_points["transform"]
    k = t.shape[1]
    diff = torch.bmm(t, t.transpose(1, 2)) - torch.eye(k, device=t.device)
    return ce + reg_weight * 0.5 * (diff ** 2).sum()
