"""Training-mode restatement of a set-abstraction level: forward with batch statistics, and the backward pass
(TEST INFRASTRUCTURE ONLY -- the checker for the next scope rows; nothing in the product imports it).

PARITY UNPINNED, like mlp_oracle.py: conv / batch-norm / relu / reduce_max and their gradients live in TensorFlow 1.x.
What is restated are the published semantics at the reference's call sites:
  pointnet2/utils/tf_util.py:155-185   conv2d 1x1 = matmul over channels + bias, then batch norm, then relu
  pointnet2/utils/tf_util.py:512-531   tf.contrib.layers.batch_norm(center, scale, is_training, decay=bn_decay,
                                       updates_collections=None): training output uses the BATCH mean and the BIASED batch
                                       variance over all axes but the channel, eps = 1e-3 (contrib default);
                                       moving_x <- decay * moving_x + (1 - decay) * batch_x   (in-place, same step)
  dgcnn/utils/tf_util.py:462-499       hand-written variant: tf.nn.moments (biased) + tf.train.ExponentialMovingAverage
                                       on the two TENSORS (shadow starts at 0, zero-debiased average), eps = 1e-3
  pointnet2/utils/pointnet_util.py:22-56,113-127   sample_and_group (indices carry no gradient: ops.NoGradient,
                                       tf_sampling.py:23,58 / tf_grouping.py:22,33), MLP, reduce_max over nsample
  tf_grouping.py:43-47 / tf_grouping_g.cu:61-78    GroupPointGrad = scatter-add of the incoming gradient by idx
  tf_sampling.py:44-48 / tf_sampling_g.cu:183-192  GatherPointGrad = scatter-add by idx
The backward formulas below are checked against central finite differences of the forward in float64
(tests/test_train_oracle.py), which is what "correct" means until a TF build is available."""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3


# ------------------------------------------------------------------------------------------------
# layers: forward returns (output, cache), backward returns gradients w.r.t. inputs and parameters
# ------------------------------------------------------------------------------------------------
def conv1x1_fwd(x, w, b):
    """x (..., Cin), w (Cin, Cout), b (Cout)"""
    return x @ w + b, (x, w)


def conv1x1_bwd(dy, cache):
    x, w = cache
    x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, dy.shape[-1])
    return (dy @ w.T), x2.T @ dy2, dy2.sum(0)


def bn_train_fwd(y, gamma, beta, eps=BN_EPS):
    """Batch statistics over every axis but the last; biased variance (tf.nn.moments / contrib batch_norm)."""
    red = tuple(range(y.ndim - 1))
    mean = y.mean(axis=red)
    var = ((y - mean) ** 2).mean(axis=red)
    inv = 1.0 / np.sqrt(var + eps)
    xhat = (y - mean) * inv
    return xhat * gamma + beta, (xhat, inv, gamma, red), mean, var


def bn_train_bwd(dz, cache):
    xhat, inv, gamma, red = cache
    m = np.prod([xhat.shape[a] for a in red])
    dgamma = (dz * xhat).sum(axis=red)
    dbeta = dz.sum(axis=red)
    dxhat = dz * gamma
    dy = (inv / m) * (m * dxhat - dxhat.sum(axis=red) - xhat * (dxhat * xhat).sum(axis=red))
    return dy, dgamma, dbeta


def moving_average_contrib(moving, batch, decay):
    """tf.contrib.layers.batch_norm update (assign_moving_average, zero_debias=False)."""
    return decay * moving + (1.0 - decay) * batch


def moving_average_ema_tensor(shadow, biased_acc, local_step, batch, decay):
    """tf.train.ExponentialMovingAverage.apply on a Tensor (dgcnn's template): zero-debiased.  Returns the new
    (average, biased accumulator, local_step); initial state (0, 0, 0)."""
    biased_acc = decay * biased_acc + (1.0 - decay) * batch
    local_step = local_step + 1
    return biased_acc / (1.0 - decay ** local_step), biased_acc, local_step


def relu_fwd(z):
    return np.maximum(z, 0), z > 0


def maxpool_fwd(h, axis=2):
    """reduce_max over the nsample axis; ties route the gradient to the FIRST maximum (argmax), one winner per cell.
    (TF's reduce_max gradient splits it equally among ties; ties have measure zero for real activations except at
    relu's 0 plateau, where the incoming relu mask is 0 anyway.)"""
    idx = np.argmax(h, axis=axis)
    return np.take_along_axis(h, np.expand_dims(idx, axis), axis).squeeze(axis), (idx, h.shape, axis)


def maxpool_bwd(dp, cache):
    idx, shape, axis = cache
    dh = np.zeros(shape, dtype=dp.dtype)
    np.put_along_axis(dh, np.expand_dims(idx, axis), np.expand_dims(dp, axis), axis)
    return dh


def group_fwd(points, idx):
    """group_point: points (B,n,C), idx (B,m,K) -> (B,m,K,C)"""
    b = np.arange(points.shape[0])[:, None, None]
    return points[b, idx]


def group_bwd(dgrouped, idx, n):
    """GroupPointGrad: scatter-add (tf_grouping_g.cu:61-78)"""
    bsz, m, k, c = dgrouped.shape
    out = np.zeros((bsz, n, c), dtype=dgrouped.dtype)
    for b in range(bsz):
        np.add.at(out[b], idx[b].reshape(-1), dgrouped[b].reshape(-1, c))
    return out


# ------------------------------------------------------------------------------------------------
# one set-abstraction level, training mode
# ------------------------------------------------------------------------------------------------
def sa_level_train_fwd(xyz, points, new_xyz_idx, idx, layers):
    """xyz (B,n,3), points (B,n,C) or None, new_xyz_idx (B,m) = FPS indices, idx (B,m,K) = ball-query indices (both
    gradient-free), layers = [(W, b, gamma, beta), ...].  -> pooled (B,m,C_L), cache, [(batch_mean, batch_var), ...]"""
    b = np.arange(xyz.shape[0])[:, None]
    new_xyz = xyz[b, new_xyz_idx]                                   # gather_point
    grouped_xyz = group_fwd(xyz, idx) - new_xyz[:, :, None, :]
    h = grouped_xyz if points is None else np.concatenate([grouped_xyz, group_fwd(points, idx)], axis=-1)
    caches, stats = [], []
    for (w, bias, gamma, beta) in layers:
        y, c_conv = conv1x1_fwd(h, w, bias)
        z, c_bn, mean, var = bn_train_fwd(y, gamma, beta)
        h, mask = relu_fwd(z)
        caches.append((c_conv, c_bn, mask))
        stats.append((mean, var))
    pooled, c_pool = maxpool_fwd(h, axis=2)
    return pooled, (caches, c_pool, idx, new_xyz_idx, xyz.shape, None if points is None else points.shape), stats


def sa_level_train_bwd(dpooled, cache):
    """-> dxyz (B,n,3), dpoints (B,n,C) or None, [(dW, db, dgamma, dbeta), ...]"""
    caches, c_pool, idx, new_xyz_idx, xyz_shape, pts_shape = cache
    dh = maxpool_bwd(dpooled, c_pool)
    grads = []
    for (c_conv, c_bn, mask) in reversed(caches):
        dz = dh * mask
        dy, dgamma, dbeta = bn_train_bwd(dz, c_bn)
        dh, dw, db = conv1x1_bwd(dy, c_conv)
        grads.append((dw, db, dgamma, dbeta))
    grads.reverse()
    n = xyz_shape[1]
    dgx = dh[..., :3]
    dxyz = group_bwd(dgx, idx, n)                                   # through grouped_xyz
    dnew = -dgx.sum(axis=2)                                         # through "- new_xyz" (B,m,3)
    for b in range(xyz_shape[0]):
        np.add.at(dxyz[b], new_xyz_idx[b], dnew[b])                 # GatherPointGrad
    dpoints = None if pts_shape is None else group_bwd(dh[..., 3:], idx, n)
    return dxyz, dpoints, grads
