import torch, sys
sys.path.insert(0, ".")
from scanobjectnn_b200 import ops
from scanobjectnn_b200.synthetic import make_clouds
x = torch.from_numpy(make_clouds("ball", 32, 2048, seed=5)).cuda()
_, q = ops.farthest_point_sample_and_gather(512, x)
for _ in range(3): ops.knn_point(32, x, q)
for k in (1, 4, 16, 31, 32):
    ts=[]
    for _ in range(5):
        torch.cuda._sleep(200000)
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); ops.knn_point(k, x, q); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    print("knn_point k=%d: %.1f us" % (k, sorted(ts)[2]))
