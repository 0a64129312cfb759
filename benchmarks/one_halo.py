"""Where does the time of the halo-slab / s2 / march kernels go?  Timing probes (no MMAs / no stores) at the KITTI shapes.
    python benchmarks/one_halo.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dcanet_b200 as d

E, L = d.engine, d._lib
bn = lambda c: torch.nn.BatchNorm3d(c).cuda().eval()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


cases = []
for name, ci, co, mode, shape in (("halo 32->32 @1/8", 32, 32, E.K3S1, (24, 48, 156)), ("halo 64->64 @1/8", 64, 64, E.K3S1, (24, 48, 156)),
                                  ("s2slab 32->64 @1/4", 32, 64, E.K3S2, (48, 96, 312)), ("march 32->32 @1/4", 32, 32, E.K3S1, (48, 96, 312))):
    x = E.Planes(1, *shape, ci, 2, "cuda"); x.t.normal_()
    pc = E.PackedConv(torch.randn(co, ci, 3, 3, 3, device="cuda") * 0.05, bn(co)); pc.pack_tc(2)
    cases.append((name, x, pc, mode))
for name, x, pc, mode in cases:
    row = []
    for flags in (0, 1, 2, 3):
        L.call("dca_tc_set_tuning", 1, flags << 4)
        row.append(timeit(lambda: E.conv(x, pc, mode, E.ACT_RELU)))
    L.call("dca_tc_set_tuning", 1, 0)
    print(f"{name:22s} full {row[0]:6.1f}  no-stores {row[1]:6.1f}  no-MMA {row[2]:6.1f}  neither {row[3]:6.1f} us")
