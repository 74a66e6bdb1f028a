"""The float64 restatement of the classifier's training step (oracle/train_model_oracle.py) against central finite
differences of its own loss: what "correct gradients" means while no TensorFlow build is available (CPU only)."""
import numpy as np

from oracle import train_model_oracle as M
from scanobjectnn_b200.synthetic import make_clouds

LEVELS = [("layer1", 12, 0.7, 5, [8, 8], False), ("layer2", None, None, None, [16], True)]
HEAD = [("fc1", 8, True, 0.5), ("fc3", None, False, None)]


def _params(rng, num_class=4):
    p = {}

    def conv(scope, cin, cout, bn=True):
        p[f"{scope}/weights"] = rng.standard_normal((cin, cout)) * 0.5
        p[f"{scope}/biases"] = rng.standard_normal(cout) * 0.1
        if bn:
            p[f"{scope}/bn/gamma"] = rng.uniform(0.5, 1.5, cout)
            p[f"{scope}/bn/beta"] = rng.standard_normal(cout) * 0.1
    conv("layer1/conv0", 3, 8); conv("layer1/conv1", 8, 8)
    conv("layer2/conv0", 3 + 8, 16)
    conv("fc1", 16, 8); conv("fc3", 8, num_class, bn=False)
    return p


def test_model_gradients_match_finite_differences():
    rng = np.random.default_rng(3)
    xyz = make_clouds("ball", 6, 40, seed=9)
    labels = rng.integers(0, 4, 6)
    p = _params(rng)
    masks = {"fc1": (rng.uniform(size=(6, 8)) < 0.5) / 0.5}
    out = M.cls_train_step(xyz, labels, p, LEVELS, HEAD, masks, 4)
    assert np.isfinite(out["loss"])
    eps = 1e-6
    checked = 0
    for name in ("layer1/conv0/weights", "layer1/conv1/bn/gamma", "layer2/conv0/weights", "layer2/conv0/bn/beta", "fc1/weights", "fc3/biases",
                 "layer1/conv1/weights", "fc1/bn/gamma"):
        g = out["grads"][name]
        flat = p[name].reshape(-1)
        for e in rng.choice(flat.size, size=min(4, flat.size), replace=False):
            old = flat[e]
            flat[e] = old + eps
            lp = M.cls_train_step(xyz, labels, p, LEVELS, HEAD, masks, 4)["loss"]
            flat[e] = old - eps
            lm = M.cls_train_step(xyz, labels, p, LEVELS, HEAD, masks, 4)["loss"]
            flat[e] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - g.reshape(-1)[e]) < 1e-6 + 1e-5 * abs(fd), (name, e, fd, g.reshape(-1)[e])
            checked += 1
    assert checked >= 24
    # a conv / fc bias in front of batch norm has a zero gradient (sum_r dy = 0)
    assert np.abs(out["grads"]["layer1/conv0/biases"]).max() < 1e-12
    assert np.abs(out["grads"]["fc1/biases"]).max() < 1e-12


def test_adam_update_first_steps():
    p, m, v = np.array([1.0, -2.0]), np.zeros(2), np.zeros(2)
    g = np.array([0.5, -0.25])
    p1, m1, v1 = M.adam_update(p, g, m, v, 1, 1e-3)
    # first Adam step moves every coordinate by ~lr against the gradient sign
    np.testing.assert_allclose(p1, p - 1e-3 * np.sign(g), rtol=0, atol=1e-7)
    p2, m2, v2 = M.adam_update(p1, g, m1, v1, 2, 1e-3)
    np.testing.assert_allclose(p2, p1 - 1e-3 * np.sign(g), rtol=0, atol=1e-7)
