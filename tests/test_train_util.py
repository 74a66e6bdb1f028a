"""Schedules and vote aggregation of the reference scripts (scanobjectnn_b200/train_util.py), CPU."""
import numpy as np

from scanobjectnn_b200 import train_util as tu


def test_learning_rate_schedule():
    # train.py:116-124 with the script defaults: 0.001 * 0.7 ** floor(batch*32 / 200000), floor 1e-5
    assert tu.get_learning_rate(0, 32) == 0.001
    assert tu.get_learning_rate(6249, 32) == 0.001                       # 199968 < 200000
    assert abs(tu.get_learning_rate(6250, 32) - 0.0007) < 1e-18          # exactly one decay step
    assert abs(tu.get_learning_rate(6250 * 3, 32) - 0.001 * 0.7 ** 3) < 1e-18
    assert tu.get_learning_rate(6250 * 40, 32) == 0.00001                # clipped


def test_bn_decay_schedule():
    # train.py:126-134: 1 - 0.5 * 0.5 ** floor(...), capped at 0.99
    assert tu.get_bn_decay(0, 32) == 0.5
    assert tu.get_bn_decay(6250, 32) == 0.75
    assert tu.get_bn_decay(6250 * 2, 32) == 0.875
    assert tu.get_bn_decay(6250 * 10, 32) == 0.99


def test_vote_aggregation():
    ang = tu.vote_angles(4)
    assert np.allclose(ang, [0, np.pi / 2, np.pi, 3 * np.pi / 2])
    votes = [np.array([[0.1, 0.9], [0.6, 0.4]]), np.array([[0.8, 0.2], [0.3, 0.7]]), np.array([[0.3, 0.6], [0.2, 0.8]])]
    pred, counts = tu.aggregate_votes(votes)
    assert pred.tolist() == [1, 1]                                       # summed scores: [1.2, 1.7], [1.1, 1.9]
    assert counts.tolist() == [[1, 2], [1, 2]]
