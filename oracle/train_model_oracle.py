"""Training step of the PointNet++ classifier restated in float64 numpy (TEST INFRASTRUCTURE ONLY -- nothing in the
product imports it).  Composes oracle/train_oracle.py's level forward/backward (checked against central finite differences
in tests/test_train_oracle.py) with the FC head, dropout, the loss and tf.train.AdamOptimizer's update.

PARITY UNPINNED like the rest of the floating-point half: the arithmetic of these ops lives in TensorFlow 1.x.  Restated
call sites: pointnet2/models/pointnet2_cls_ssg.py:23-57 (levels, fc1/dp1/fc2/dp2/fc3, mean sparse softmax
cross-entropy), pointnet2/utils/tf_util.py:187-229 (fully_connected = matmul + bias, batch norm, relu), :533-548
(dropout = tf.nn.dropout: keep w.p. keep_prob, scale by 1/keep_prob), pointnet2/train.py:139-146 (Adam, default
beta1=0.9, beta2=0.999, epsilon=1e-8; update lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps))."""
from __future__ import annotations

import numpy as np

from . import oracle as orc
from . import train_oracle as T


def _layer_params(params, scope):
    w = np.asarray(params[f"{scope}/weights"], dtype=np.float64)
    w = w.reshape(-1, w.shape[-1])
    b = np.asarray(params[f"{scope}/biases"], dtype=np.float64)
    if f"{scope}/bn/gamma" in params:
        return w, b, np.asarray(params[f"{scope}/bn/gamma"], np.float64), np.asarray(params[f"{scope}/bn/beta"], np.float64)
    return w, b, None, None


def cls_train_step(xyz, labels, params, levels, head, masks, num_class):
    """xyz (B,N,3) float32, labels (B,) int, params {tf name: array}, levels = [(scope, npoint, radius, nsample, mlp, group_all)],
    head = [(scope, width, bn, keep)], masks = {scope: (B,width) dropout mask of 0 / 1/keep}.
    -> dict(logits, loss, grads {name: array}, batch_stats {scope: (mean, var)}, idx [per level])"""
    xyz32 = np.asarray(xyz, dtype=np.float32)
    B = xyz32.shape[0]
    cur_xyz32, cur_pts = xyz32, None
    caches, all_stats, idxs = [], {}, []
    for (scope, npoint, radius, nsample, mlp, group_all) in levels:
        layers = [_layer_params(params, f"{scope}/conv{i}") for i in range(len(mlp))]
        if group_all:
            n = cur_xyz32.shape[1]
            fps_idx = np.zeros((B, 1), dtype=np.int64)
            idx = np.broadcast_to(np.arange(n)[None, None, :], (B, 1, n))
            # sample_and_group_all: new_xyz = 0 (pointnet_util.py:70), so grouped_xyz - new_xyz = xyz
            x = cur_xyz32.astype(np.float64)
            h = x[:, None, :, :] if cur_pts is None else np.concatenate([x, cur_pts], axis=-1)[:, None, :, :]
            lc, stats = [], []
            for (w, b, gamma, beta) in layers:
                y, c_conv = T.conv1x1_fwd(h, w, b)
                z, c_bn, mean, var = T.bn_train_fwd(y, gamma, beta)
                h, mask = T.relu_fwd(z)
                lc.append((c_conv, c_bn, mask))
                stats.append((mean, var))
            pooled, c_pool = T.maxpool_fwd(h, axis=2)
            caches.append(("all", lc, c_pool, None if cur_pts is None else cur_pts.shape[-1]))
            new_xyz32 = np.zeros((B, 1, 3), dtype=np.float32)
        else:
            fps_idx = orc.fps(cur_xyz32, npoint)
            new_xyz32 = orc.gather_point(cur_xyz32, fps_idx)
            idx, _ = orc.query_ball_point(radius, nsample, cur_xyz32, new_xyz32, contract=True)
            # float32 centring exactly as the kernels / the reference do it (grouped_xyz - new_xyz in fp32), then float64
            pooled, cache, stats = T.sa_level_train_fwd(cur_xyz32, cur_pts, fps_idx.astype(np.int64), idx.astype(np.int64), layers)
            caches.append(("sa", cache))
        for i, st in enumerate(stats):
            all_stats[f"{scope}/conv{i}"] = st
        idxs.append(idx)
        cur_xyz32, cur_pts = new_xyz32, pooled
    feat = cur_pts.reshape(B, -1)
    h = feat
    hc = []
    for (scope, width, bn, keep) in head:
        w, b, gamma, beta = _layer_params(params, scope)
        y, c_conv = T.conv1x1_fwd(h, w, b)
        if bn:
            z, c_bn, mean, var = T.bn_train_fwd(y, gamma, beta)
            all_stats[scope] = (mean, var)
            h, rmask = T.relu_fwd(z)
        else:
            c_bn, rmask, h = None, None, y
        dm = masks.get(scope) if keep is not None else None
        if dm is not None:
            h = h * dm
        hc.append((c_conv, c_bn, rmask, dm))
    logits = h
    # mean sparse softmax cross-entropy
    mx = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - mx)
    sm = e / e.sum(axis=1, keepdims=True)
    lab = np.asarray(labels).astype(np.int64)
    loss = float(np.mean(np.log(e.sum(axis=1)) + mx[:, 0] - logits[np.arange(B), lab]))
    dlog = sm.copy()
    dlog[np.arange(B), lab] -= 1.0
    dlog /= B
    grads = {}
    dh = dlog
    for (scope, width, bn, keep), (c_conv, c_bn, rmask, dm) in zip(reversed(head), reversed(hc)):
        if dm is not None:
            dh = dh * dm
        if bn:
            dz = dh * rmask
            dy, dgamma, dbeta = T.bn_train_bwd(dz, c_bn)
            grads[f"{scope}/bn/gamma"], grads[f"{scope}/bn/beta"] = dgamma, dbeta
        else:
            dy = dh
        dh, dw, db = T.conv1x1_bwd(dy, c_conv)
        grads[f"{scope}/weights"], grads[f"{scope}/biases"] = dw, db
    dpooled = dh.reshape(B, 1, -1)
    for (scope, npoint, radius, nsample, mlp, group_all), cache in zip(reversed(levels), reversed(caches)):
        if cache[0] == "all":
            _, lc, c_pool, cfeat = cache
            d = T.maxpool_bwd(dpooled, c_pool)
            lg = []
            for (c_conv, c_bn, mask) in reversed(lc):
                dz = d * mask
                dy, dgamma, dbeta = T.bn_train_bwd(dz, c_bn)
                d, dw, db = T.conv1x1_bwd(dy, c_conv)
                lg.append((dw, db, dgamma, dbeta))
            lg.reverse()
            dpts = None if cfeat is None else d[:, 0, :, 3:]
        else:
            _, dpts, lg = T.sa_level_train_bwd(dpooled, cache[1])
        for i, (dw, db, dgamma, dbeta) in enumerate(lg):
            s = f"{scope}/conv{i}"
            grads[f"{s}/weights"], grads[f"{s}/biases"] = dw, db
            grads[f"{s}/bn/gamma"], grads[f"{s}/bn/beta"] = dgamma, dbeta
        dpooled = dpts
    return dict(logits=logits, loss=loss, grads=grads, batch_stats=all_stats, idx=idxs)


def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer._apply_dense: returns (p, m, v) after step `step` (1-based)."""
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    lr_t = lr * np.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v
