#!/usr/bin/env python
"""cva.classify.2's shifted sum + class statistics at the KITTI 1/8-res shape: one fused launch vs the two kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
E = d.engine
Ps = [torch.randn(27, 1, 24, 48, 156, device="cuda") for _ in range(2)]


def t(fn, n=50):
    for _ in range(5):
        fn(0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(n):
        fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for _ in range(2):
    print("fused                         %6.1f us" % t(lambda i: E.tap_gather_class_stats(Ps[i % 2])))
    print("tap_gather3d + class_stats    %6.1f us" % t(lambda i: E.class_stats(E.tap_gather(Ps[i % 2]))))
    print("tap_gather3d alone            %6.1f us" % t(lambda i: E.tap_gather(Ps[i % 2])))
