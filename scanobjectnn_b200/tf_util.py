"""The slice of the reference's tf_util.py files that sits on the hot path, on torch CUDA tensors.

* variable store with the reference's TF variable names (``layer1/conv0/weights``, ``.../bn/gamma`` ...), so a
  TF checkpoint name map is the identity;
* ``conv2d`` 1x1 / ``fully_connected`` (+bias +batch norm +ReLU) in inference mode, executed by the hand-written
  dense-layer kernel (pointnet2/utils/tf_util.py:120-185, 512-531; dgcnn/utils/tf_util.py:115-173, 462-499);
* DGCNN's graph functions (dgcnn/utils/tf_util.py:638-706).
"""
from __future__ import annotations

import math

import torch

from . import ops
from .ops import get_edge_feature, knn, knn_graph, pairwise_distance  # noqa: F401

BN_EPS = 1e-3   # tf.contrib.layers.batch_norm default (pointnet2 tf_util.py:526-531); explicit in dgcnn tf_util.py:498


class VariableStore(dict):
    """name -> tensor, named exactly as the reference's TF variable scopes name them."""

    def __init__(self, device="cuda", seed: int = 0):
        super().__init__()
        self.device = torch.device(device)
        self._gen = torch.Generator(device="cpu")
        self._gen.manual_seed(seed)
        self._cache = {}

    # --- initialisers (tf_util._variable_with_weight_decay with use_xavier=True; biases 0; BN gamma 1 / beta 0) ---
    def _xavier(self, shape, fan_in, fan_out):
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        w = (torch.rand(shape, generator=self._gen, dtype=torch.float32) * 2 - 1) * limit
        return w.to(self.device)

    def _bn(self, scope, c, randomize):
        if randomize:   # non-trivial statistics so parity tests exercise the folding
            r = lambda lo, hi: (torch.rand(c, generator=self._gen) * (hi - lo) + lo).to(self.device)
            self[f"{scope}/bn/beta"] = r(-0.1, 0.1)
            self[f"{scope}/bn/gamma"] = r(0.8, 1.2)
            self[f"{scope}/bn/moving_mean"] = r(-0.1, 0.1)
            self[f"{scope}/bn/moving_variance"] = r(0.5, 1.5)
        else:
            self[f"{scope}/bn/beta"] = torch.zeros(c, device=self.device)
            self[f"{scope}/bn/gamma"] = torch.ones(c, device=self.device)
            self[f"{scope}/bn/moving_mean"] = torch.zeros(c, device=self.device)
            self[f"{scope}/bn/moving_variance"] = torch.ones(c, device=self.device)

    def add_conv2d(self, scope, cin, cout, bn=True, randomize_bn=False):
        self[f"{scope}/weights"] = self._xavier((1, 1, cin, cout), cin, cout)
        self[f"{scope}/biases"] = torch.zeros(cout, device=self.device)
        if bn:
            self._bn(scope, cout, randomize_bn)

    def add_fc(self, scope, cin, cout, bn=True, randomize_bn=False):
        self[f"{scope}/weights"] = self._xavier((cin, cout), cin, cout)
        self[f"{scope}/biases"] = torch.zeros(cout, device=self.device)
        if bn:
            self._bn(scope, cout, randomize_bn)

    # --- inference-mode folding: y = relu((x.W) * scale + shift) ---
    def folded(self, scope, relu=True):
        """(W (Cin,Cout), scale or None, shift, relu) for ``scope``; BN folded if ``scope/bn/gamma`` exists."""
        key = ("layer", scope, relu)
        if key not in self._cache:
            w = self[f"{scope}/weights"]
            w2 = w.reshape(-1, w.shape[-1]).contiguous().float()
            b = self[f"{scope}/biases"].float()
            if f"{scope}/bn/gamma" in self:
                inv = self[f"{scope}/bn/gamma"].float() * torch.rsqrt(self[f"{scope}/bn/moving_variance"].float() + BN_EPS)
                shift = (b - self[f"{scope}/bn/moving_mean"].float()) * inv + self[f"{scope}/bn/beta"].float()
                self._cache[key] = (w2, inv.contiguous(), shift.contiguous(), relu)
            else:
                self._cache[key] = (w2, None, b.contiguous(), relu)
        return self._cache[key]

    def mlp(self, scopes, relus=None, xyz_last: bool = False) -> ops.MlpParams:
        """``xyz_last``: the first layer's input is [features, xyz] (pointnet_sa_module_msg concatenates that way,
        pointnet_util.py:184) -- its last three weight rows are rotated to the front, which is where the fused kernels
        expect the coordinate rows."""
        relus = relus if relus is not None else [True] * len(scopes)
        key = ("mlp", tuple(scopes), tuple(relus), xyz_last)
        if key not in self._cache:
            layers = [self.folded(s, r) for s, r in zip(scopes, relus)]
            if xyz_last:
                w, sc, sh, r = layers[0]
                layers[0] = (torch.cat([w[-3:], w[:-3]], dim=0).contiguous(), sc, sh, r)
            self._cache[key] = ops.MlpParams(layers)
        return self._cache[key]

    def invalidate(self):
        """Drop the folded conv+BN tensors / prepared weight images.  Called automatically when a variable is assigned,
        updated or deleted; call it yourself after an IN-PLACE edit of a weight tensor.  Inference engines / CUDA graphs
        captured earlier keep the old images: rebuild them after a weight change."""
        self._cache.clear()

    # assigning variables (e.g. loading a checkpoint after a forward) must not leave stale folded weights behind
    def __setitem__(self, key, value):
        if getattr(self, "_cache", None):
            self._cache.clear()
        super().__setitem__(key, value)

    def __delitem__(self, key):
        if getattr(self, "_cache", None):
            self._cache.clear()
        super().__delitem__(key)

    def update(self, *args, **kwargs):
        if getattr(self, "_cache", None):
            self._cache.clear()
        super().update(*args, **kwargs)


def _require_inference(is_training):
    if is_training:
        raise NotImplementedError(
            "is_training=True is not available for this configuration; training mode covers the models' own layer configurations "
            "(training.py: sa_module_training, mlp_training, PointNet2ClsTrainer) -- see INTEGRATION.md")


def _layer_training(inputs, scope, activation_fn, bn, bn_decay, params):
    """one conv / fully-connected layer in training mode (training.mlp_training): conv+BN+ReLU or a plain linear layer"""
    if bn and activation_fn is not None:
        kind = True
    elif not bn and activation_fn is None:
        kind = False
    else:
        raise NotImplementedError("training mode covers conv/fc + batch norm + ReLU and plain linear layers (what the models use)")
    from .training import mlp_training
    return mlp_training(inputs, [(scope, kind)], bn_decay, params)


def conv2d(inputs, num_output_channels, kernel_size, scope, stride=(1, 1), padding="SAME", data_format="NHWC",
           activation_fn="relu", bn=False, bn_decay=None, is_training=False, *, params: VariableStore):
    """tf_util.conv2d restricted to what the hot path uses: 1x1 kernels, stride 1, NHWC, ReLU or None."""
    if tuple(kernel_size) != (1, 1) or tuple(stride) != (1, 1) or data_format != "NHWC":
        raise NotImplementedError("only 1x1 / stride-1 / NHWC convolutions are on the point-set-abstraction path")
    if is_training:
        return _layer_training(inputs, scope, activation_fn, bn, bn_decay, params)
    relu = activation_fn is not None
    mlp = params.mlp([scope], [relu])
    if mlp.channels[-1] != num_output_channels:
        raise ValueError(f"{scope}: stored weights have {mlp.channels[-1]} outputs, asked for {num_output_channels}")
    return ops.shared_mlp(inputs, mlp)


def fully_connected(inputs, num_outputs, scope, activation_fn="relu", bn=False, bn_decay=None, is_training=False, *,
                    params: VariableStore):
    if is_training:
        return _layer_training(inputs, scope, activation_fn, bn, bn_decay, params)
    relu = activation_fn is not None
    mlp = params.mlp([scope], [relu])
    if mlp.channels[-1] != num_outputs:
        raise ValueError(f"{scope}: stored weights have {mlp.channels[-1]} outputs, asked for {num_outputs}")
    return ops.shared_mlp(inputs, mlp)


def dropout(inputs, is_training, scope, keep_prob=0.5, noise_shape=None):
    """tf_util.dropout (tf_util.py:560-580): identity at inference, tf.nn.dropout(keep_prob) in training"""
    if is_training:
        if noise_shape is not None:
            raise NotImplementedError("dropout: noise_shape is not used by the in-scope models")
        return torch.nn.functional.dropout(inputs, 1.0 - keep_prob, training=True)
    return inputs
