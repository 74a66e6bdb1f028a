"""The public inference API (scanobjectnn_b200.engine): several batches in flight on several streams, CUDA-graph
replay, pinned-host in/out -- results must equal the plain eager forward for every batch."""
import numpy as np
import pytest
import torch

from scanobjectnn_b200 import pointnet2_cls_ssg
from scanobjectnn_b200.engine import pointnet2_cls_ssg_engine
from scanobjectnn_b200.synthetic import make_clouds

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("slots", [1, 3])
def test_engine_matches_eager_forward(slots):
    params = pointnet2_cls_ssg.init_params(seed=4, randomize_bn=True)
    eng = pointnet2_cls_ssg_engine(params, batch=4, npoints=1024, slots=slots)
    batches = [torch.from_numpy(make_clouds(kind, 4, 1024, seed=50 + i)) for i, kind in enumerate(["ball", "shell", "dup", "ball", "shell"])]
    want = [pointnet2_cls_ssg.get_model(b.cuda(), False, params=params)[0].cpu().numpy() for b in batches]
    # device-resident submission
    got = []
    for b in batches:
        slot = eng.submit(b.cuda())
        got.append(eng.result(slot).cpu().numpy().copy())
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # pinned-host submission with all slots in flight before any result is read
    pinned = [b.pin_memory() for b in batches[:slots]]
    slots_used = [eng.submit(p, to_host=True) for p in pinned]
    for s_, w in zip(slots_used, want[:slots]):
        assert np.array_equal(eng.result(s_, host=True).numpy(), w)
