#!/usr/bin/env python
"""1/8-res 32->32 3x3x3 conv (cva.downsample / cva.classify.0 shape at KITTI): halo-slab kernel vs the depth-marching kernel
with different planes-per-item, plain and with the fused 32->1 tail epilogue."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
E, L = d.engine, d._lib
dims = tuple(int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (1, 24, 48, 156)
B, D, H, W = dims
x = E.Planes(B, D, H, W, 32, 2, "cuda"); x.t.normal_()
bn = torch.nn.BatchNorm3d(32).cuda().eval()
pc = E.PackedConv(torch.randn(32, 32, 3, 3, 3, device="cuda") * 0.05, bn); pc.pack_tc(2)
pc1 = E.PackedCout1(torch.randn(1, 32, 3, 3, 3, device="cuda") * 0.05, 2)


def t(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


E.Options.march_min_items = 10 ** 9
ref = E.conv(x, pc, E.K3S1, E.ACT_RELU).t.clone()
print("halo kernel            %6.1f us   taps27 %6.1f us" % (t(lambda: E.conv(x, pc, E.K3S1, E.ACT_RELU)), t(lambda: E.conv_taps27(x, pc, pc1))))
E.Options.march_min_items = 1
for n in (0, 8, 7, 6, 5, 4, 3):
    L.call("dca_tc_set_march_n", n)
    y = E.conv(x, pc, E.K3S1, E.ACT_RELU).t
    print("march n=%d (0 = auto)   %6.1f us   taps27 %6.1f us   max diff vs halo %.2e" % (
        n, t(lambda: E.conv(x, pc, E.K3S1, E.ACT_RELU)), t(lambda: E.conv_taps27(x, pc, pc1)),
        float((y.float() - ref.float()).abs().max())))
L.call("dca_tc_set_march_n", 0)
