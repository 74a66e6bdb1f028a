"""Seeded synthetic point clouds shaped like what the reference's training scripts feed the models
after ``center_data`` + ``normalize_data`` (data_utils.py:133-168): (B,N,3) float32, zero-mean,
max-norm exactly 1.  Generators follow SURVEY.md 8(d):

  ball   uniform in the unit ball            -- worst case for ball query (few early exits)
  shell  noisy unit-sphere surface (s=0.01)  -- scan-like, early exits exercised
  dup    ball with 25 % of the points overwritten by point 0 (tie-breaking; what
         provider.random_point_dropout does, pointnet2/utils/provider.py:229-236)
"""
from __future__ import annotations

import numpy as np


def center_normalize(pc: np.ndarray) -> np.ndarray:
    """data_utils.center_data + normalize_data (data_utils.py:133-168), per cloud."""
    pc = pc.astype(np.float32)
    pc = pc - pc.mean(axis=1, keepdims=True, dtype=np.float32)
    scale = np.sqrt((pc.astype(np.float32) ** 2).sum(-1)).max(axis=1)
    return (pc / scale[:, None, None]).astype(np.float32)


def make_clouds(kind: str, b: int, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind in ("ball", "dup"):
        v = rng.standard_normal((b, n, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        r = rng.random((b, n, 1)) ** (1.0 / 3.0)
        pc = v * r
        if kind == "dup":
            mask = rng.random((b, n)) < 0.25
            pc = np.where(mask[..., None], pc[:, :1, :], pc)
    elif kind == "shell":
        v = rng.standard_normal((b, n, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        pc = v + 0.01 * rng.standard_normal((b, n, 3))
    else:
        raise ValueError(kind)
    return center_normalize(pc.astype(np.float32))


def make_labels(b: int, num_class: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, num_class, size=(b,), dtype=np.int32)
