import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scanobjectnn_b200 import ops
from scanobjectnn_b200.synthetic import make_clouds
x = torch.from_numpy(make_clouds("ball", 32, 2048, seed=1001)).cuda()
f = torch.randn((32, 2048, 64), device="cuda")
for _ in range(2):
    ops.knn_graph(x, 20); ops.knn_graph(f, 20)
torch.cuda.synchronize(); print("ok")
