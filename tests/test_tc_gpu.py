"""tcgen05 path: one tensor-core layer in isolation (descriptor / TMEM layout / swizzle check), then the fused
set-abstraction kernel in both arithmetic modes against fp64."""
import numpy as np
import pytest
import torch

from scanobjectnn_b200 import ops

from . import gpu_util as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kd,n", [(64, 64), (64, 128), (128, 64), (128, 128)])
def test_tc_single_layer_matches_fp64(kd, n):
    rng = np.random.default_rng(kd * 1000 + n)
    a = np.maximum(rng.standard_normal((128, kd)), 0).astype(np.float32)
    a[5] = 0.0
    a[:, 3] = np.arange(128, dtype=np.float32) / 128         # a structured column/row pattern: catches layout mix-ups
    w = (rng.uniform(-1, 1, (kd, n)) * np.sqrt(6.0 / (kd + n))).astype(np.float32)
    w[7, :] = np.linspace(-1, 1, n, dtype=np.float32)
    d = G.npy(ops.tc_selftest(G.cu(a), G.cu(w)))
    want = a.astype(np.float64) @ w.astype(np.float64)
    err = np.abs(d - want).max()
    print(f"tc layer {kd}x{n}: max|err|={err:.3e} max|d|={np.abs(want).max():.3f}")
    assert err < 1e-5 * max(1.0, np.abs(want).max()), err


def test_tc_identity_weight_is_exact_passthrough():
    a = np.random.default_rng(0).standard_normal((128, 128)).astype(np.float32)
    d = G.npy(ops.tc_selftest(G.cu(a), G.cu(np.eye(128, dtype=np.float32))))
    # w = 1 is exact in tf32, so d = trunc(a) + tf32(a - trunc(a)) + 0: within 2^-21 relative of a
    assert np.abs(d - a).max() <= np.abs(a).max() * 2.0 ** -20


def test_mlp_mode_switch():
    assert ops.get_mlp_mode() == 0
    ops.set_mlp_mode(1)
    assert ops.get_mlp_mode() == 1
    ops.set_mlp_mode(0)
    with pytest.raises(ValueError):
        ops.set_mlp_mode(7)
