"""Individual hot ops at the bench shape, for ncu: FPS, ball query, F1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scanobjectnn_b200 import ops, pointnet2_cls_ssg
from scanobjectnn_b200.synthetic import make_clouds
p = pointnet2_cls_ssg.init_params(seed=1, randomize_bn=True)
x = torch.from_numpy(make_clouds("ball", 32, 2048, seed=1001)).cuda()
for _ in range(2):
    _, l1 = ops.farthest_point_sample_and_gather(512, x)
    ops.query_ball_point(0.2, 32, x, l1)
    _, l2 = ops.farthest_point_sample_and_gather(128, l1)
    ops.query_ball_point(0.4, 64, l1, l2)
    ops.sa_conv1_prebn(x, l1, None, 0.2, 32, p["layer1/conv0/weights"].reshape(3, 64), p["layer1/conv0/biases"], want_stats=True)
torch.cuda.synchronize()
print("ok")
