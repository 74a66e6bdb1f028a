"""Model assemblies against the fp64 restatement: pointnet2_cls_bga end to end; dgcnn / dgcnn_bga
stage by stage (each stage's oracle is fed the GPU's input features of that stage, so a last-ulp difference in a
feature cannot flip a near-tie neighbour and turn a 1e-6 error into a different graph)."""
import numpy as np
import pytest
import torch

from oracle import mlp_oracle as mo
from oracle import oracle as orc
from scanobjectnn_b200 import dgcnn, ops, pointnet2_cls_bga, pointnet_cls
from scanobjectnn_b200.synthetic import make_clouds

from . import gpu_util as G

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _close(got, want, tol=TOL, what=""):
    if tol == TOL:
        return G.contract_close(got, want, what)        # 1e-5 absolute up to |act| = 1, 1e-5 * max|act| above (printed)
    err = float(np.abs(got - want).max())
    scale = max(1.0, float(np.abs(want).max()))
    assert err < tol * scale, (what, err, scale)
    return err


@pytest.mark.parametrize("kind", ["ball", "shell"])
def test_pointnet2_cls_bga_matches_oracle(kind):
    p = pointnet2_cls_bga.init_params(seed=3, randomize_bn=True)
    xyz = make_clouds(kind, 2, 2048, seed=4001)
    cls, seg, ep = pointnet2_cls_bga.get_model(G.cu(xyz), False, params=p, return_end_points=True)
    assert len(pointnet2_cls_bga.get_model(G.cu(xyz), False, params=p)) == 2      # the reference's (class_pred, seg_pred)
    wcls, wseg = mo.pointnet2_cls_bga(xyz, p)
    assert seg.shape == (2, 2048, 2) and cls.shape == (2, 15)
    e1 = _close(G.npy(cls), wcls, what="class_pred")
    e2 = _close(G.npy(seg), wseg, tol=2e-5, what="seg_pred")
    print(f"bga[{kind}] class err {e1:.2e} seg err {e2:.2e}")
    loss, cl, sl = pointnet2_cls_bga.get_loss(cls, seg, torch.zeros(2, dtype=torch.int64, device="cuda"),
                                              torch.zeros((2, 2048), dtype=torch.int64, device="cuda"))
    assert torch.isfinite(loss) and abs(float(loss) - 0.5 * float(cl) - 0.5 * float(sl)) < 1e-6


def test_fp_module_with_single_known_point_broadcasts():
    # fa_layer1 of BGA: three_nn against ONE known point -> dist (d,inf,inf), idx (0,0,0), weights (1,0,0)
    xyz1 = G.cu(make_clouds("ball", 2, 128, seed=5))
    xyz2 = torch.zeros((2, 1, 3), device="cuda")
    pts2 = torch.randn((2, 1, 16), device="cuda")
    out, dist, idx, w = ops.three_nn_interpolate(xyz1, xyz2, pts2, return_aux=True)
    assert torch.equal(out, pts2.expand(2, 128, 16))
    assert torch.isinf(dist[..., 1:]).all() and (idx == 0).all()
    assert torch.equal(w, torch.tensor([1.0, 0.0, 0.0], device="cuda").expand(2, 128, 3))


@pytest.mark.parametrize("bga,n", [(False, 512), (True, 512), (False, 2048)])      # 2048 = BASELINE.json configs[2]
def test_dgcnn_stagewise_matches_oracle(bga, n):
    p = dgcnn.init_params(seed=5, randomize_bn=True, bga=bga)
    # a non-trivial input transform (the reference initialises it to identity)
    p["transform_net1/transform_XYZ/weights"] = 0.01 * torch.randn((256, 9), device="cuda")
    xyz = make_clouds("ball", 2, n, seed=4002)
    x = G.cu(xyz)
    if bga:
        cls, seg, ep = dgcnn.get_model_bga(x, False, params=p, return_end_points=True)
        assert seg.shape == (2, n, 2)
    else:
        cls, ep = dgcnn.get_model(x, False, params=p)
    assert cls.shape == (2, 15) and torch.isfinite(cls).all()
    # stage 0: graph on the raw cloud + T-net
    assert np.array_equal(G.npy(ep["nn_idx0"]), orc.dgcnn_knn(xyz, 20))
    t = mo.edgeconv(xyz, G.npy(ep["nn_idx0"]), p, ["transform_net1/tconv1", "transform_net1/tconv2"])
    t = mo.mlp_chain(t, p, ["transform_net1/tconv3"]).max(axis=1)
    t = mo.mlp_chain(t, p, ["transform_net1/tfc1", "transform_net1/tfc2"])
    tr = t @ G.npy(p["transform_net1/transform_XYZ/weights"]).astype(np.float64) + np.eye(3).flatten()
    _close(G.npy(ep["transform"]).reshape(2, 9), tr, what="transform")
    # stages 1-4: teacher-forced on the GPU's stage input
    feats = G.npy(ep["point_cloud_transformed"])
    for i, scope in enumerate(["dgcnn1", "dgcnn2", "dgcnn3", "dgcnn4"]):
        idx, y = mo.dgcnn_stage(feats, 20, p, [scope])
        assert np.array_equal(G.npy(ep[f"nn_idx{i + 1}"]), idx), scope
        _close(G.npy(ep[f"net{i + 1}"]), y, what=scope)
        feats = G.npy(ep[f"net{i + 1}"])
    cat = np.concatenate([G.npy(ep[f"net{i}"]) for i in (1, 2, 3, 4)], -1)
    agg = mo.mlp_chain(cat, p, ["agg"])
    glob = agg.max(axis=1)
    if bga:
        net = mo.mlp_chain(glob, p, ["fc1", "fc2"])
        _close(G.npy(cls), mo.mlp_chain(net, p, ["fc3"], [False]), what="class_pred")
        concat = np.concatenate([np.broadcast_to(net[:, None], (2, n, 256)), np.broadcast_to(glob[:, None], (2, n, 1024)), cat], -1)
        wseg = mo.mlp_chain(concat, p, ["seg/conv1", "seg/conv2", "seg/conv3"], [True, True, False])
        _close(G.npy(seg), wseg, tol=2e-5, what="seg_pred")
    else:
        _close(G.npy(ep["global"]), glob, what="global")
        _close(G.npy(cls), mo.mlp_chain(glob, p, ["fc1", "fc2", "fc3"], [True, True, False]), what="logits")


def test_pointnet_cls_vanilla_matches_oracle():
    """BASELINE.json configs[0]: PointNet vanilla, B=8 N=1024 (the reference's plumbing case)."""
    p = pointnet_cls.init_params(seed=6, randomize_bn=True)
    for name in ("transform_net1/transform_XYZ/weights", "transform_net2/transform_feat/weights"):
        p[name] = 0.01 * torch.randn(p[name].shape, device="cuda")        # non-identity transforms
    xyz = make_clouds("shell", 8, 1024, seed=4003)
    logits, ep = pointnet_cls.get_model(G.cu(xyz), False, params=p)
    want, t2 = mo.pointnet_cls(xyz, p)
    _close(G.npy(ep["transform"]), t2, what="feature transform")
    _close(G.npy(logits), want, what="logits")
    loss = pointnet_cls.get_loss(logits, torch.zeros(8, dtype=torch.int64, device="cuda"), ep)
    assert torch.isfinite(loss)
