"""Property-based pin of the CPU restatement (oracle/psa_oracle.c) against the reference's OWN CPU code compiled here
(oracle/_ref/libref_cpu.so: grouping/test/query_ball_point.cpp, selection_sort.cpp, tf_interpolate.cpp:57-153): random shapes down
to the degenerate ones (n = 1, m = 1, nsample > n, k = n) and adversarial values -- points on a lattice so that many distances sit
exactly on the radius, duplicated points (ties), tiny and large magnitudes.  Bit-exact, CPU only."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import oracle as orc

pytestmark = pytest.mark.skipif(not orc.refcpu_available(), reason="oracle/_ref/libref_cpu.so not built")
FUZZ = settings(max_examples=200, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


def _cloud(rng, b, n, flavour):
    if flavour == 0:                                   # lattice: distances repeat and land exactly on round radii
        x = rng.integers(-4, 5, size=(b, n, 3)).astype(np.float32) * np.float32(0.125)
    elif flavour == 1:                                 # duplicates of a few points
        base = rng.standard_normal((b, max(n // 4, 1), 3)).astype(np.float32)
        x = base[:, rng.integers(0, base.shape[1], size=n), :]
    elif flavour == 2:                                 # tiny magnitudes (denormal squares)
        x = (rng.standard_normal((b, n, 3)) * 1e-20).astype(np.float32)
    elif flavour == 3:                                 # large offset: catastrophic cancellation in the differences
        x = (rng.standard_normal((b, n, 3)) * 0.1 + 1000.0).astype(np.float32)
    else:
        x = rng.standard_normal((b, n, 3)).astype(np.float32)
    return np.ascontiguousarray(x)


@FUZZ
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 3), n=st.integers(1, 70), m=st.integers(1, 20), nsample=st.integers(1, 80),
       flavour=st.integers(0, 4), radius=st.sampled_from([0.125, 0.25, 0.5, 0.2, 1e-3, 3.0, 1e-19]))
def test_ball_query_fuzz(seed, b, n, m, nsample, flavour, radius):
    rng = np.random.default_rng(seed)
    xyz = _cloud(rng, b, n, flavour)
    q = _cloud(rng, b, m, flavour) if seed % 2 else np.ascontiguousarray(xyz[:, rng.integers(0, n, size=m), :])
    oi, cnt = orc.query_ball_point(radius, nsample, xyz, q, contract=False, fill=-7)     # contract=False: the CPU harness's arithmetic
    ri = orc.refcpu_query_ball_point(radius, nsample, xyz, q, fill=-7)
    assert np.array_equal(oi, ri)
    assert ((cnt == 0) == (oi[..., 0] == -7)).all() and cnt.max() <= nsample


@FUZZ
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 3), m=st.integers(1, 6), n=st.integers(1, 40), ties=st.booleans(), kfrac=st.floats(0.0, 1.0))
def test_selection_sort_fuzz(seed, b, m, n, ties, kfrac):
    rng = np.random.default_rng(seed)
    dist = rng.random((b, m, n), dtype=np.float32)
    if ties:
        dist = np.round(dist * 4).astype(np.float32) / 4           # five distinct values: position-dependent tie resolution
    k = max(1, min(n, int(round(kfrac * n))))
    oi, ov = orc.selection_sort(k, dist)
    ri, rv = orc.refcpu_selection_sort(k, dist)
    assert np.array_equal(oi, ri) and np.array_equal(ov, rv)


@FUZZ
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 3), n=st.integers(1, 50), m=st.integers(1, 30), flavour=st.integers(0, 4))
def test_three_nn_fuzz(seed, b, n, m, flavour):
    rng = np.random.default_rng(seed)
    xyz1, xyz2 = _cloud(rng, b, n, flavour), _cloud(rng, b, m, flavour)
    od, oi = orc.three_nn(xyz1, xyz2)
    rd, ri = orc.refcpu_three_nn(xyz1, xyz2)
    assert np.array_equal(oi, ri) and np.array_equal(od, rd)


@FUZZ
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 3), m=st.integers(1, 20), n=st.integers(1, 30), c=st.integers(1, 9), dup=st.booleans())
def test_three_interpolate_and_grad_fuzz(seed, b, m, n, c, dup):
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((b, m, c)).astype(np.float32)
    idx = rng.integers(0, m, size=(b, n, 3), dtype=np.int32)
    if dup:
        idx[..., 1] = idx[..., 0]                                  # the same source twice in a row: order of the scatter adds matters
    w = rng.random((b, n, 3), dtype=np.float32)
    assert np.array_equal(orc.three_interpolate(pts, idx, w), orc.refcpu_three_interpolate(pts, idx, w))
    go = rng.standard_normal((b, n, c)).astype(np.float32)
    assert np.array_equal(orc.three_interpolate_grad(pts.shape, idx, w, go), orc.refcpu_three_interpolate_grad(pts.shape, idx, w, go))


@FUZZ
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 3), n=st.integers(1, 30), c=st.integers(1, 9), m=st.integers(1, 12), k=st.integers(1, 9))
def test_group_point_and_grad_fuzz(seed, b, n, c, m, k):
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((b, n, c)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, m, k), dtype=np.int32)
    assert np.array_equal(orc.group_point(pts, idx), orc.refcpu_group_point(pts, idx))
    go = rng.standard_normal((b, m, k, c)).astype(np.float32)
    assert np.array_equal(orc.group_point_grad(pts.shape, idx, go), orc.refcpu_group_point_grad(pts.shape, idx, go))
