"""AvgPool3d kernels at the KITTI shape: thread-per-output vs depth-marching TMA kernel with different depth splits."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dcanet_b200 as d

E, L = d.engine, d._lib
x = E.Planes(1, 48, 96, 312, 32, 2, "cuda"); x.t.normal_()
x2 = E.Planes(1, 48, 96, 312, 32, 2, "cuda"); x2.t.normal_()


def timeit(n=20):
    for _ in range(3):
        E.avgpool(x)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(n):
        E.avgpool(x if i % 2 else x2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


L.call("dca_pool_set_march", 0)
print("thread-per-output      %6.1f us" % timeit())
for c in (2, 3, 4, 5, 6, 8, 10, 12):
    L.call("dca_pool_set_march", c)
    print("march, %2d CTAs per SM   %6.1f us" % (c, timeit()))
L.call("dca_pool_set_march", 4)
