"""Eager (no CUDA graph) forwards of pointnet2_cls_ssg at the bench shape, for ncu.
usage: python tools/profile_step.py [n_iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from scanobjectnn_b200 import pointnet2_cls_ssg
from scanobjectnn_b200.synthetic import make_clouds

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
params = pointnet2_cls_ssg.init_params(seed=1, randomize_bn=True)
x = torch.from_numpy(make_clouds("ball", 32, 2048, seed=1001)).cuda()
for _ in range(n):
    logits, _ = pointnet2_cls_ssg.get_model(x, False, params=params)
torch.cuda.synchronize()
print("ok", float(logits.abs().max()))
