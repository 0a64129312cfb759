#!/usr/bin/env python
"""One launch each of the small-K tcgen05 kernels at KITTI shapes (for `ncu --set full`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
E = d.engine
P = 2
q4, q8 = (1, 48, 96, 312), (1, 24, 48, 156)
x4 = E.Planes(*q4, 32, P, "cuda"); x4.t.normal_()
r4 = E.Planes(*q4, 32, P, "cuda"); r4.t.normal_()
x8 = E.Planes(*q8, 64, P, "cuda"); x8.t.normal_()
t8 = E.Planes(1, 48, 50, 158, 32, P, "cuda"); t8.t.normal_()
bn = torch.nn.BatchNorm3d(32).cuda().eval()
pc1 = E.PackedConv(torch.randn(32, 32, 1, 1, 1, device="cuda") * .1, bn); pc1.pack_tc(P)
pcs2 = E.PackedConv(torch.randn(64, 32, 3, 3, 3, device="cuda") * .05, torch.nn.BatchNorm3d(64).cuda().eval()); pcs2.pack_tc(P)
lib = d._lib.load()
w28 = torch.zeros(28 * lib.dca_pack_weights_tc_bytes(32, 64, 1, P), dtype=torch.uint8, device="cuda")
w4 = torch.zeros(lib.dca_pack_weights_tc_bytes(32, 32, 4, P), dtype=torch.uint8, device="cuda")
sc = torch.ones(32, device="cuda"); sh = torch.zeros(32, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(n):
    E.conv(x4, pc1, E.K1, E.ACT_NONE)
    E.conv(x4, pcs2, E.K3S2, E.ACT_RELU)
    E.up2(0, x8, x4, w28, sc, sh, E.ACT_RELU, 64, 24, 48, 156, res_post=r4)
    E.up2(2, t8, x4, w4, sc, sh, E.ACT_NONE, 32, 48, 48, 156)
torch.cuda.synchronize(); print("ok")
