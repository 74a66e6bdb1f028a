"""Host-side training / evaluation utilities of the reference scripts (pure Python, no device code):
the learning-rate and batch-norm-decay schedules of pointnet2/train.py:116-134 and the rotation-vote aggregation of
pointnet2/evaluate_scenennobjects.py:170-196."""
from __future__ import annotations

import math

import numpy as np

# defaults of pointnet2/train.py (argparse defaults and module constants)
BN_INIT_DECAY = 0.5
BN_DECAY_DECAY_RATE = 0.5
BN_DECAY_CLIP = 0.99


def exponential_decay(base: float, global_step: int, decay_steps: int, decay_rate: float, staircase: bool = True) -> float:
    """tf.train.exponential_decay: base * decay_rate ** (global_step / decay_steps), the exponent floored if staircase."""
    p = global_step / float(decay_steps)
    if staircase:
        p = math.floor(p)
    return base * decay_rate ** p


def get_learning_rate(batch: int, batch_size: int, base_learning_rate: float = 0.001, decay_step: int = 200000,
                      decay_rate: float = 0.7) -> float:
    """train.py:116-124: staircase exponential decay of the index into the dataset (batch * BATCH_SIZE), clipped at 1e-5."""
    return max(exponential_decay(base_learning_rate, batch * batch_size, decay_step, decay_rate), 0.00001)


def get_bn_decay(batch: int, batch_size: int, bn_decay_decay_step: int = 200000) -> float:
    """train.py:126-134: bn_decay = min(BN_DECAY_CLIP, 1 - 0.5 * 0.5 ** floor(batch * BATCH_SIZE / step))."""
    bn_momentum = exponential_decay(BN_INIT_DECAY, batch * batch_size, bn_decay_decay_step, BN_DECAY_DECAY_RATE)
    return min(BN_DECAY_CLIP, 1 - bn_momentum)


def vote_angles(num_votes: int):
    """evaluate_scenennobjects.py:180-182: vote v rotates the batch about the up axis by v / num_votes * 2 pi."""
    return [v / float(num_votes) * np.pi * 2 for v in range(num_votes)]


def aggregate_votes(pred_vals):
    """evaluate_scenennobjects.py:178-193: pred_vals = list over votes of (B, num_classes) scores ->
    (pred (B,) = argmax of the SUMMED scores, per-class vote counts (B, num_classes) as the script also tallies)."""
    pred_vals = [np.asarray(p) for p in pred_vals]
    batch_pred_sum = np.zeros_like(pred_vals[0], dtype=np.float64)
    batch_pred_classes = np.zeros(pred_vals[0].shape)
    for p in pred_vals:
        batch_pred_sum += p
        batch_pred_classes[np.arange(p.shape[0]), np.argmax(p, 1)] += 1
    return np.argmax(batch_pred_sum, 1), batch_pred_classes
