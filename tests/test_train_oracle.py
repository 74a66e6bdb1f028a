"""The training-mode restatement (oracle/train_oracle.py): backward formulas against central finite differences of the
forward in float64, batch-norm statistics semantics, moving-average rules.  CPU only; this is the checker the fused
training kernels of the next round will be held to."""
import numpy as np

from oracle import train_oracle as to


def _level(seed=0, B=2, n=24, m=5, K=6, C=4, widths=(7, 5)):
    rng = np.random.default_rng(seed)
    xyz = rng.standard_normal((B, n, 3))
    pts = rng.standard_normal((B, n, C))
    fps = np.stack([rng.choice(n, m, replace=False) for _ in range(B)])
    idx = rng.integers(0, n, (B, m, K))
    layers, cin = [], 3 + C
    for w in widths:
        layers.append([rng.standard_normal((cin, w)) * 0.5, rng.standard_normal(w) * 0.1, rng.uniform(0.5, 1.5, w), rng.standard_normal(w) * 0.1])
        cin = w
    return xyz, pts, fps, idx, layers


def _loss(xyz, pts, fps, idx, layers, proj):
    out, _, _ = to.sa_level_train_fwd(xyz, pts, fps, idx, [tuple(l) for l in layers])
    return float((out * proj).sum())


def test_backward_matches_finite_differences():
    xyz, pts, fps, idx, layers = _level()
    rng = np.random.default_rng(9)
    out, cache, stats = to.sa_level_train_fwd(xyz, pts, fps, idx, [tuple(l) for l in layers])
    proj = rng.standard_normal(out.shape)
    dxyz, dpts, grads = to.sa_level_train_bwd(proj, cache)
    eps = 1e-6

    def fd(arr, pick):
        old = arr[pick]
        arr[pick] = old + eps; lp = _loss(xyz, pts, fps, idx, layers, proj)
        arr[pick] = old - eps; lm = _loss(xyz, pts, fps, idx, layers, proj)
        arr[pick] = old
        return (lp - lm) / (2 * eps)

    for pick in [(0, 3, 1), (1, 7, 2), (0, int(fps[0, 0]), 0)]:
        assert abs(fd(xyz, pick) - dxyz[pick]) < 1e-5 * max(1.0, abs(dxyz[pick])), ("xyz", pick)
    for pick in [(0, int(idx[0, 0, 0]), 1), (1, int(idx[1, 2, 3]), 3)]:
        assert abs(fd(pts, pick) - dpts[pick]) < 1e-5 * max(1.0, abs(dpts[pick])), ("points", pick)
    for li, (dw, db, dg, dbeta) in enumerate(grads):
        for arr, g, pick in [(layers[li][0], dw, (1, 2)), (layers[li][1], db, (0,)), (layers[li][2], dg, (3,)), (layers[li][3], dbeta, (1,))]:
            assert abs(fd(arr, pick) - g[pick]) < 1e-5 * max(1.0, abs(g[pick])), (li, pick)
    # the conv bias in front of a training-mode batch norm has zero gradient (the batch mean removes it)
    assert all(np.abs(g[1]).max() < 1e-9 for g in grads)


def test_batch_norm_statistics_and_moving_averages():
    rng = np.random.default_rng(1)
    y = rng.standard_normal((3, 4, 5, 6)) * 2 + 1
    z, _, mean, var = to.bn_train_fwd(y, np.ones(6), np.zeros(6))
    flat = y.reshape(-1, 6)
    assert np.allclose(mean, flat.mean(0)) and np.allclose(var, flat.var(0))          # biased variance
    assert np.allclose(z.reshape(-1, 6).mean(0), 0, atol=1e-12)
    assert np.allclose(z.reshape(-1, 6).var(0), var / (var + to.BN_EPS))
    mv = to.moving_average_contrib(np.ones(6), var, 0.9)
    assert np.allclose(mv, 0.9 + 0.1 * var)
    avg, acc, step = to.moving_average_ema_tensor(0.0, 0.0, 0, mean, 0.9)
    assert np.allclose(avg, mean) and step == 1                                          # zero-debiased: first average = first value
    avg2, _, _ = to.moving_average_ema_tensor(avg, acc, step, mean * 3, 0.9)
    assert np.allclose(avg2, (0.9 * 0.1 * mean + 0.1 * 3 * mean) / (1 - 0.81))


def test_group_and_pool_routing():
    rng = np.random.default_rng(2)
    pts = rng.standard_normal((2, 9, 3))
    idx = rng.integers(0, 9, (2, 4, 5))
    g = to.group_fwd(pts, idx)
    assert g.shape == (2, 4, 5, 3) and np.array_equal(g[1, 2, 3], pts[1, idx[1, 2, 3]])
    dg = rng.standard_normal(g.shape)
    back = to.group_bwd(dg, idx, 9)
    want = np.zeros_like(pts)
    for b in range(2):
        for i in range(4):
            for j in range(5):
                want[b, idx[b, i, j]] += dg[b, i, j]
    assert np.allclose(back, want)
    h = rng.standard_normal((2, 4, 5, 3))
    p, c = to.maxpool_fwd(h)
    assert np.array_equal(p, h.max(2))
    dh = to.maxpool_bwd(np.ones_like(p), c)
    assert dh.sum() == p.size and np.array_equal(dh.argmax(2), h.argmax(2))
