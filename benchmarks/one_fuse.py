"""Time dca_up2_tc kind 2 (bilinear x2 + cat + 1x1x1 fuse + BN, cva.py:64,55,69) at the KITTI shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dcanet_b200 as d
import workloads

E, L = d.engine, d._lib
net = workloads.init_bench_weights_(d.GwcNet(192), 0).cuda().eval()
pk = net.packed().cva[0]
B, D8, H8, W8 = 1, 24, 48, 156
t = E.Planes(B, 2 * D8, H8 + 2, W8 + 2, 32, 2, "cuda"); t.t.normal_()
costs = [E.Planes(B, 2 * D8, 2 * H8, 2 * W8, 32, 2, "cuda") for _ in range(2)]
for c in costs:
    c.t.normal_()


def run(n=20):
    f = lambda i: E.up2(2, t, costs[i % 2], pk.fuse_up2b_w, pk.fuse_scale, pk.fuse_shift, E.ACT_NONE, 32, 2 * D8, H8, W8)
    for i in range(3):
        f(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(n):
        f(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for slots in (2, 4):
    L.call("dca_tc_set_up2_side_slots", slots)
    row = []
    for flags in (0, 1, 2, 3):
        L.call("dca_tc_set_tuning", 1, flags << 4)
        row.append(run())
    L.call("dca_tc_set_tuning", 1, 0)
    print(f"side slots {slots}: full {row[0]:6.1f}  no-stores {row[1]:6.1f}  no-MMA {row[2]:6.1f}  neither {row[3]:6.1f} us")
L.call("dca_tc_set_up2_side_slots", 2)
