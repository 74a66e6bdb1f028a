"""Input pipeline: ops.augment_batch against the numpy restatement of data_utils / provider."""
import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as po
from scanobjectnn_b200 import ops

from . import gpu_util as G

pytestmark = pytest.mark.gpu


def _raw(b, n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((b, n, 3)) * np.array([1.5, 0.7, 1.1]) + np.array([0.3, -2.0, 0.9])).astype(np.float32)


def test_rotate_jitter_is_bit_exact():
    """train.py:246-247: rotate_point_cloud then jitter_point_cloud -- float64 products and sums, stored as float32."""
    b, n = 16, 2048
    rng = np.random.default_rng(5)
    x = po.normalize_data(po.center_data(_raw(b, n, 1)))
    angles = rng.uniform(size=b) * 2 * np.pi
    noise = rng.standard_normal((b, n, 3)).astype(np.float32)
    got = G.npy(ops.augment_batch(G.cu(x), angles=angles, noise=G.cu(noise)))
    want = po.augment(x, n, angles=angles, noise=noise)
    # cos/sin come from torch (CPU, float64) vs numpy: identical libm results in practice; allow 1 ulp of float32 for the few
    # elements where the float64 product lands on a rounding boundary
    assert np.abs(got - want).max() <= 2.4e-7 * max(1.0, np.abs(want).max())
    assert (got != want).mean() < 1e-3


def test_center_normalize_subset():
    b, n_src, n = 8, 2500, 1024
    x = _raw(b, n_src, 2)
    perm = np.random.default_rng(3).permutation(n_src).astype(np.int32)
    got = G.npy(ops.augment_batch(G.cu(x), n, perm=G.cu(perm), center=True, normalize=True))
    want = po.augment(x, n, perm=perm, center=True, normalize=True)
    assert got.shape == (b, n, 3)
    assert np.abs(got - want).max() < 2e-6          # the float32 centroid sums are ordered differently
    full = G.npy(ops.augment_batch(G.cu(x), center=True, normalize=True))
    assert abs(np.sqrt((full ** 2).sum(-1)).max(1) - 1).max() < 1e-6 and np.abs(full.mean(1)).max() < 1e-5


def test_all_steps_and_errors():
    b, n = 4, 300
    rng = np.random.default_rng(9)
    x = _raw(b, n, 4)
    kw = dict(angles=rng.uniform(size=b) * 6.28, scale=rng.uniform(0.8, 1.25, b).astype(np.float32),
              shift=rng.uniform(-0.1, 0.1, (b, 3)).astype(np.float32), noise=rng.standard_normal((b, n, 3)).astype(np.float32),
              drop=rng.random((b, n)) < 0.3)
    got = G.npy(ops.augment_batch(G.cu(x), **{k: (G.cu(v) if k not in ("angles",) else v) for k, v in kw.items()}))
    want = po.augment(x, n, **kw)
    assert np.abs(got - want).max() <= 5e-7 * max(1.0, np.abs(want).max())
    with pytest.raises(ValueError):
        ops.augment_batch(G.cu(x), n + 1)
    with pytest.raises(ValueError):
        ops.augment_batch(G.cu(x), noise=G.cu(kw["noise"]), clip=0.0)
    d = ops.draw_augmentation(b, n, 128, "cuda", torch.Generator(device="cuda").manual_seed(1))
    out = ops.augment_batch(G.cu(x), 128, **d)
    assert out.shape == (b, 128, 3) and torch.isfinite(out).all()
