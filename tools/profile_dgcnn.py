"""Eager (no CUDA graph) DGCNN forwards at the bench shape (B=32, N=2048, k=20), for an ncu launch list.
usage: python tools/profile_dgcnn.py [n_iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from scanobjectnn_b200 import dgcnn
from scanobjectnn_b200.synthetic import make_clouds

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
params = dgcnn.init_params(seed=3, randomize_bn=True)
x = torch.from_numpy(make_clouds("ball", 32, 2048, seed=1001)).cuda()
for _ in range(n):
    out = dgcnn.get_model(x, False, params=params)[0]
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
