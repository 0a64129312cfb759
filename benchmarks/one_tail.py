#!/usr/bin/env python
"""dca_tap_gather_softmax_regress at the KITTI shape under its disparity-group count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
E, L = d.engine, d._lib
B, D, H, W = 1, 48, 96, 312
Ps = [torch.randn(27, B, D, H, W, device="cuda") for _ in range(2)]
ref = None
for n in (16, 24, 32, 16, 24, 32):
    L.call("dca_tap_gather_set_groups", n)
    for _ in range(3):
        out = E.tap_gather_softmax_regress(Ps[0])[0]
    if ref is None:
        ref = out.clone()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(20):
        E.tap_gather_softmax_regress(Ps[i % 2])
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    print(f"groups {n:2d}: {us:6.1f} us  {Ps[0].numel() * 4 / us / 1e3:.0f} GB/s  max diff vs 16 groups {float((out - ref).abs().max()):.2e}")
L.call("dca_tap_gather_set_groups", 16)
